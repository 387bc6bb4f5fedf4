// Packed genome-list streams on the device: cursor (reader) and writer.
//
// Layout (maple_b200/genome_list.py): one uint32 key per entry
//   bits 0-2 type | 3-4 nLens | 5 flag | 6-7 nuc | 8-31 end (1-based inclusive)
// and a float64 payload stream holding, per entry, nLens branch lengths followed (type 6) by
// the 4-vector.  A reader never needs the payload of R/N runs it skips, so the cursor only
// advances a payload pointer and loads lengths / vectors on demand at the few informative sites.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace maple {

constexpr int T_R = 4, T_N = 5, T_O = 6;

// LD = true: read-only arena data through the non-coherent path (__ldg);
// LD = false: scratch written earlier by the same thread (plain loads).
template <bool LD>
struct Cursor {
    const uint32_t* key;
    const double* pay;   // payload of the CURRENT entry
    int type, nl, flag, nuc, end;

    __device__ __forceinline__ void decode(uint32_t k) {
        type = int(k & 7u);
        nl = int((k >> 3) & 3u);
        flag = int((k >> 5) & 1u);
        nuc = int((k >> 6) & 3u);
        end = int(k >> 8);
    }
    __device__ __forceinline__ uint32_t ldk(const uint32_t* p) const { return LD ? __ldg(p) : *p; }
    __device__ __forceinline__ double ldp(const double* p) const { return LD ? __ldg(p) : *p; }

    __device__ __forceinline__ void init(const uint32_t* k, const double* p) {
        key = k;
        pay = p;
        decode(ldk(key));
    }
    __device__ __forceinline__ void next() {
        pay += nl + (type == T_O ? 4 : 0);
        ++key;
        decode(ldk(key));
    }
    __device__ __forceinline__ double l0() const { return nl >= 1 ? ldp(pay) : 0.0; }
    __device__ __forceinline__ double l1() const { return nl == 2 ? ldp(pay + 1) : 0.0; }
    __device__ __forceinline__ double v(int i) const { return ldp(pay + nl + i); }
    __device__ __forceinline__ void vec(double* o) const {
        const double* q = pay + nl;
        o[0] = ldp(q); o[1] = ldp(q + 1); o[2] = ldp(q + 2); o[3] = ldp(q + 3);
    }
};

struct Writer {
    uint32_t* key;
    double* pay;
    int nk, np;
    __device__ __forceinline__ void init(uint32_t* k, double* p) { key = k; pay = p; nk = 0; np = 0; }
    __device__ __forceinline__ void put(int type, int nl, int flag, int nuc, int end, double l0, double l1, const double* vec) {
        key[nk++] = uint32_t(type & 7) | (uint32_t(nl & 3) << 3) | (uint32_t(flag ? 1 : 0) << 5) | (uint32_t(nuc & 3) << 6) |
                    (uint32_t(end) << 8);
        if (nl >= 1) pay[np++] = l0;
        if (nl == 2) pay[np++] = l1;
        if (type == T_O) {
            pay[np++] = vec[0]; pay[np++] = vec[1]; pay[np++] = vec[2]; pay[np++] = vec[3];
        }
    }
    __device__ __forceinline__ void put0(int type, int nuc, int end) { put(type, 0, 0, nuc, end, 0.0, 0.0, nullptr); }
};

}  // namespace maple
