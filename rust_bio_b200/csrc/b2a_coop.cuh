// Warp-cooperative primitives shared by the banded kernels (b2a_banded.cuh) and the warp-per-pair walk
// (b2a_walk.cuh): shuffles, ballots and reductions over the W lanes that work on one pair.  W = 32 is a warp on
// the device (and 32 cooperatively scheduled host contexts in tests/sim); W = 1 is the plain sequential build.
#pragma once
#include "b2a_common.cuh"

namespace b2a {

// Cooperative-lane helpers: W = 32 lanes of one warp on the device, W = 1 in the host logic build.
template <int W>
struct Coop {
  static B2A_HD void sync() {
#if defined(__CUDA_ARCH__)
    if (W > 1) __syncwarp();
#elif defined(B2A_HOST_WARP)
    if (W > 1) host_warp_sync();
#endif
  }
  static B2A_HD int32_t up(int32_t v, int d) {  // the value held by lane - d
#if defined(__CUDA_ARCH__)
    if (W > 1) return __shfl_up_sync(0xffffffffu, v, d);
#elif defined(B2A_HOST_WARP)
    if (W > 1) {
      const int me = host_lane;
      return (int32_t)host_warp_exchange(v, [&](const long long* x) { return me >= d ? x[me - d] : x[me]; });
    }
#endif
    (void)d;
    return v;
  }
  static B2A_HD int32_t from(int32_t v, int src) {
#if defined(__CUDA_ARCH__)
    if (W > 1) return __shfl_sync(0xffffffffu, v, src);
#elif defined(B2A_HOST_WARP)
    if (W > 1) return (int32_t)host_warp_exchange(v, [&](const long long* x) { return x[src & 31]; });
#endif
    (void)src;
    return v;
  }
  static B2A_HD uint32_t ballot(bool b) {  // bit l = lane l's predicate
#if defined(__CUDA_ARCH__)
    if (W > 1) return __ballot_sync(0xffffffffu, b);
#elif defined(B2A_HOST_WARP)
    if (W > 1)
      return (uint32_t)host_warp_exchange(b ? 1 : 0, [&](const long long* x) {
        long long m = 0;
        for (int l = 0; l < 32; ++l) m |= (x[l] ? 1ll : 0ll) << l;
        return m;
      });
#endif
    return b ? 1u : 0u;
  }
  static B2A_HD unsigned long long all_sum(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    if (W > 1)
      for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
#elif defined(B2A_HOST_WARP)
    if (W > 1)
      return (unsigned long long)host_warp_exchange((long long)v, [&](const long long* x) {
        unsigned long long t = 0;
        for (int l = 0; l < 32; ++l) t += (unsigned long long)x[l];
        return (long long)t;
      });
#endif
    return v;
  }
  // claim an empty (zero) 64-bit slot: returns true if `val` was stored
  static B2A_HD bool claim(uint64_t* slot, uint64_t val) {
#if defined(__CUDA_ARCH__)
    if (W > 1)
      return atomicCAS(reinterpret_cast<unsigned long long*>(slot), 0ull, (unsigned long long)val) == 0ull;
#elif defined(B2A_HOST_WARP)
    if (W > 1) {
      uint64_t expect = 0;
      return __atomic_compare_exchange_n(slot, &expect, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    }
#endif
    if (*slot != 0) return false;
    *slot = val;
    return true;
  }
  static B2A_HD uint32_t fetch_add(uint32_t* ctr, uint32_t v) {
#if defined(__CUDA_ARCH__)
    if (W > 1) return atomicAdd(ctr, v);
#elif defined(B2A_HOST_WARP)
    if (W > 1) return __atomic_fetch_add(ctr, v, __ATOMIC_SEQ_CST);
#endif
    const uint32_t old = *ctr;
    *ctr = old + v;
    return old;
  }
  static B2A_HD void or_u32(uint32_t* p, uint32_t v) {  // set bits in a word other lanes may be setting too
#if defined(__CUDA_ARCH__)
    if (W > 1) {
      atomicOr(p, v);
      return;
    }
#elif defined(B2A_HOST_WARP)
    if (W > 1) {
      __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST);
      return;
    }
#endif
    *p |= v;
  }
  static B2A_HD int popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
  }
  static B2A_HD int32_t all_max32(int32_t v) {  // one REDUX on the device
#if defined(__CUDA_ARCH__)
    if (W > 1) return __reduce_max_sync(0xffffffffu, v);
#elif defined(B2A_HOST_WARP)
    if (W > 1)
      return (int32_t)host_warp_exchange(v, [&](const long long* x) {
        long long t = x[0];
        for (int l = 1; l < 32; ++l) t = x[l] > t ? x[l] : t;
        return t;
      });
#endif
    return v;
  }
  static B2A_HD long long all_max(long long v) {
#if defined(__CUDA_ARCH__)
    if (W > 1)
      for (int d = 16; d; d >>= 1) {
        const long long t = __shfl_xor_sync(0xffffffffu, v, d);
        v = t > v ? t : v;
      }
#elif defined(B2A_HOST_WARP)
    if (W > 1)
      return host_warp_exchange(v, [&](const long long* x) {
        long long t = x[0];
        for (int l = 1; l < 32; ++l) t = x[l] > t ? x[l] : t;
        return t;
      });
#endif
    return v;
  }
};

}  // namespace b2a
