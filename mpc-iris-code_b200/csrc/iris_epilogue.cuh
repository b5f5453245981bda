// Shared epilogue helper of the scan / batched kernels: result rows are 62 bytes (31 x u16, the reference's wire
// format, src/main.rs:429-431), so a tile's rows are staged in shared memory with the same 16-byte phase as their
// destination and then stored with 128-bit accesses.
#pragma once
#include <stdint.h>

namespace iris {

// Result stores are streaming stores (st.global.cs): the rows are written once and read much later by another kernel or
// a copy engine, and plain stores cost the HBM-bound scan 2.4 % (3.91 -> 3.81 ms per 1 M rows fused; tests/diagnostics/
// store_mode_bench.py).  IRIS_STORE_MODE (compile-time, A/B builds only): 0 = plain st.global, 1 = st.global.cs,
// 2 = st.global with an L2 evict_first policy (same time as 1), 3 = st.global with an L2 evict_last policy
// (3.87 ms against 3.82 for mode 1 and 3.91 for mode 0 on one box).  Also measured and dropped: whole tiles leaving through
// the TMA engine (cp.async.bulk shared -> global): same time as mode 1; all tiles storing into the same 0.5 MB (no DRAM
// writes, timing only): 3.78 against 3.81 ms -- the stores cost about 1 % of the fused scan.
#ifndef IRIS_STORE_MODE
#define IRIS_STORE_MODE 1
#endif
__device__ __forceinline__ void store_out16(uint8_t* g, const uint4& v) {
#if IRIS_STORE_MODE == 1
    __stcs(reinterpret_cast<uint4*>(g), v);
#elif IRIS_STORE_MODE == 2
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(g), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
#elif IRIS_STORE_MODE == 3
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(g), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
#else
    *reinterpret_cast<uint4*>(g) = v;
#endif
}

// Copies bytes [b0,b1) (offsets inside `stage`, both even) to gbase + offset, where gbase is 16-byte aligned and
// congruent with `stage`: 16-byte body, 2-byte head / tail.  Executed by the 128 epilogue threads (tid 0..127).
__device__ __forceinline__ void copy_out_rows(const uint8_t* stage, uint8_t* gbase, int b0, int b1, int tid) {
    if (b1 <= b0) return;
    int body0 = (b0 + 15) & ~15, body1 = b1 & ~15;
    if (body0 > body1) {  // shorter than one aligned vector
        for (int b = b0 + 2 * tid; b < b1; b += 2 * 128)
            *reinterpret_cast<uint16_t*>(gbase + b) = *reinterpret_cast<const uint16_t*>(stage + b);
        return;
    }
    for (int b = b0 + 2 * tid; b < body0; b += 2 * 128)
        *reinterpret_cast<uint16_t*>(gbase + b) = *reinterpret_cast<const uint16_t*>(stage + b);
    for (int b = body0 + 16 * tid; b < body1; b += 16 * 128)
        store_out16(gbase + b, *reinterpret_cast<const uint4*>(stage + b));
    for (int b = body1 + 2 * tid; b < b1; b += 2 * 128)
        *reinterpret_cast<uint16_t*>(gbase + b) = *reinterpret_cast<const uint16_t*>(stage + b);
}

}  // namespace iris
