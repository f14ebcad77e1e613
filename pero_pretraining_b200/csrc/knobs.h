// Tuning knobs.  The production build (plain `make`) has NONE: tuning_knob() is the compiled-in default, the library
// reads no environment variable and keeps no mutable global state (include/pero_b200.h, "Conventions").
// `make DEV=1` defines PERO_DEV_BUILD: every knob can then be overridden through the environment variable of the same
// name (read once per process) for A/B measurements on the GPU box, and the per-launch timeline hooks are compiled in.
#pragma once
#include <stdlib.h>

namespace pero {

#ifdef PERO_DEV_BUILD
inline int tuning_knob_lookup(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
// one static per call site (the macro expands to a lambda with its own static)
#define PERO_KNOB(name, dflt) ([]() -> int { static const int v = ::pero::tuning_knob_lookup(name, dflt); return v; }())
#else
#define PERO_KNOB(name, dflt) (dflt)
#endif

}  // namespace pero
