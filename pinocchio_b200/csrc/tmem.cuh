// tmem.cuh — Blackwell tensor memory (TMEM, 256 KB per SM) used as per-thread scratch.
//
// The tree sweeps are bound by how much per-configuration state fits on chip (shared memory: 227 KB per
// SM).  TMEM doubles that capacity: 512 columns x 128 lanes x 32 bit per SM, reached with
// tcgen05.st / tcgen05.ld (SASS STTM / LDTM, 12-cycle load latency).  With the 32x32b shape thread t of a
// warp reads / writes lane (32 * (warp % 4) + t), so a warp-uniform column index addresses "slot k of
// every thread of the warp" — exactly the access pattern of the slot-major shared state (tree.cuh).  No
// tensor-core instruction is involved; TMEM is only storage here.
//
// Protocol: one warp allocates (tcgen05.alloc) and later frees (tcgen05.dealloc) the CTA's columns; a
// store is made visible to a later load of the same thread by tcgen05.wait::st; a load's registers may be
// read after tcgen05.wait::ld.
#pragma once

#include <stdint.h>

#include "spatial.cuh"

namespace brbd
{

#define BRBD_TMEM_LD_ASM(N, ...) asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #N ".b32 " __VA_ARGS__)

BRBD_DI void tmem_ld_w(uint32_t a, uint32_t (&r)[1])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(a) : "memory");
}
BRBD_DI void tmem_ld_w(uint32_t a, uint32_t (&r)[2])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
}
BRBD_DI void tmem_ld_w(uint32_t a, uint32_t (&r)[4])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
}
BRBD_DI void tmem_ld_w(uint32_t a, uint32_t (&r)[8])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(a) : "memory");
}
BRBD_DI void tmem_ld_w(uint32_t a, uint32_t (&r)[16])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(a) : "memory");
}
BRBD_DI void tmem_st_w(uint32_t a, const uint32_t (&r)[1])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a), "r"(r[0]) : "memory");
}
BRBD_DI void tmem_st_w(uint32_t a, const uint32_t (&r)[2])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(r[0]), "r"(r[1]) : "memory");
}
BRBD_DI void tmem_st_w(uint32_t a, const uint32_t (&r)[4])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
BRBD_DI void tmem_st_w(uint32_t a, const uint32_t (&r)[8])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
BRBD_DI void tmem_st_w(uint32_t a, const uint32_t (&r)[16])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(a),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
BRBD_DI void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
BRBD_DI void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// W 32-bit words starting at column `a`, split into power-of-two chunks (largest first)
template<int W> struct TmemChunk { static constexpr int value = W >= 16 ? 16 : (W >= 8 ? 8 : (W >= 4 ? 4 : (W >= 2 ? 2 : 1))); };
template<int W> BRBD_DI void tmem_ld_words(uint32_t a, uint32_t * w)
{
  if constexpr (W > 0)
  {
    constexpr int C = TmemChunk<W>::value;
    uint32_t r[C];
    tmem_ld_w(a, r);
#pragma unroll
    for (int k = 0; k < C; ++k) w[k] = r[k];
    tmem_ld_words<W - C>(a + C, w + C);
  }
}
template<int W> BRBD_DI void tmem_st_words(uint32_t a, const uint32_t * w)
{
  if constexpr (W > 0)
  {
    constexpr int C = TmemChunk<W>::value;
    uint32_t r[C];
#pragma unroll
    for (int k = 0; k < C; ++k) r[k] = w[k];
    tmem_st_w(a, r);
    tmem_st_words<W - C>(a + C, w + C);
  }
}

// N values of type T (double = 2 columns each, float = 1) at value-offset `o` of this warp's TMEM slice.
template<class T> struct TmemSlots
{
  uint32_t base; // (lane base << 16) | first column of this warp
  static constexpr int WPV = (int)(sizeof(T) / 4);
  template<int N> BRBD_DI void load(int o, T * v) const
  {
    uint32_t w[N * WPV];
    tmem_ld_words<N * WPV>(base + (uint32_t)(o * WPV), w);
    tmem_wait_ld();
#pragma unroll
    for (int k = 0; k < N; ++k)
    {
      if constexpr (sizeof(T) == 8) v[k] = __hiloint2double((int)w[2 * k + 1], (int)w[2 * k]);
      else v[k] = __uint_as_float(w[k]);
    }
  }
  template<int N> BRBD_DI void store(int o, const T * v) const
  {
    uint32_t w[N * WPV];
#pragma unroll
    for (int k = 0; k < N; ++k)
    {
      if constexpr (sizeof(T) == 8) { w[2 * k] = (uint32_t)__double2loint(v[k]); w[2 * k + 1] = (uint32_t)__double2hiint(v[k]); }
      else w[k] = __float_as_uint(v[k]);
    }
    tmem_st_words<N * WPV>(base + (uint32_t)(o * WPV), w);
  }
};

// One warp allocates `cols` (power of two, 32..512) columns for the CTA; every thread gets the base.
BRBD_DI uint32_t tmem_alloc_cta(int cols, uint32_t * smem_slot)
{
  if ((threadIdx.x >> 5) == 0)
  {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem_slot);
    switch (cols)
    {
    case 32: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(sa) : "memory"); break;
    case 64: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(sa) : "memory"); break;
    case 128: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(sa) : "memory"); break;
    case 256: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(sa) : "memory"); break;
    default: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sa) : "memory"); break;
    }
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *smem_slot;
}
BRBD_DI void tmem_free_cta(uint32_t base, int cols)
{
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0)
  {
    switch (cols)
    {
    case 32: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base) : "memory"); break;
    case 64: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base) : "memory"); break;
    case 128: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(base) : "memory"); break;
    case 256: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base) : "memory"); break;
    default: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory"); break;
    }
  }
}
inline int tmem_round_cols(int cols) { int c = 32; while (c < cols) c <<= 1; return c; }

} // namespace brbd
