// Byte layouts of the opaque device blobs and workspaces of the C ABI (host-only helpers).
// Everything is carved out of caller-owned memory at 256-byte aligned offsets.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace pero {

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Prepared codebook: bf16 operand [K, Dp] (Dp = D rounded up to 64, zero padded) | |c|^2 fp32 [Kp]
// (Kp = K rounded up to 256, +inf padded so that out-of-range columns never win the arg-min).
struct CodebookLayout { int64_t Dp, Kp; size_t cb_off, cnorm_off, total; };
inline CodebookLayout codebook_layout(int64_t K, int64_t D) {
    CodebookLayout l;
    l.Dp = round_up(D, 64); l.Kp = round_up(K, 256);
    l.cb_off = 0;
    l.cnorm_off = align256((size_t)K * l.Dp * 2);
    l.total = l.cnorm_off + align256((size_t)l.Kp * 4);
    return l;
}

// Assign workspace: bf16 frames [N, Dp] | packed (distance, index) u64 [N].
struct AssignWsLayout { int64_t Dp; size_t xb_off, packed_off, total; };
inline AssignWsLayout assign_ws_layout(int64_t N, int64_t D) {
    AssignWsLayout l;
    l.Dp = round_up(D, 64);
    l.xb_off = 0;
    l.packed_off = align256((size_t)N * l.Dp * 2);
    l.total = l.packed_off + align256((size_t)N * 8);
    return l;
}

}  // namespace pero
