// Error codes of the C ABI (see include/pero_b200.h).  0 = success, negative = argument / environment
// error detected by the library, positive = a cudaError_t returned by the runtime.
#pragma once
#define PERO_OK 0
#define PERO_ERR_BAD_SHAPE (-1)
#define PERO_ERR_BAD_ALIGN (-2)
#define PERO_ERR_WORKSPACE (-3)
#define PERO_ERR_ARCH (-4)
#define PERO_ERR_NULL (-5)
#define PERO_ERR_DRIVER (-6)
#define PERO_ERR_UNSUPPORTED (-7)
