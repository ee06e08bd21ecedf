// fp64 pairwise contraction on the 5th-generation tensor cores (tcgen05).
//
// tcgen05.mma has no f64 kind, so large real-fp64 contractions are run as an
// ERROR-FREE integer decomposition (Ozaki scheme) on kind::i8:
//
//   A[m,k] = 2^eA[m] * sum_p qA_p[m,k] 2^-(6+7p),   qA_p in [-64, 64] (int8)
//   B[k,n] = 2^eB[n] * sum_q qB_q[k,n] 2^-(6+7q)
//   C[m,n] = 2^(eA[m]+eB[n]-12) * sum_d 2^-7d * sum_{p+q=d} sum_k qA_p qB_q
//
// Every int8 x int8 product and every int32 accumulation is exact (|sum| <
// 2^31 for (d+1) K 2^12 < 2^31, i.e. K < 65536), so the only error is the
// truncation of the operands to 6+7(S-1) bits below each row/column maximum
// and the dropped cross terms p+q >= S:  S = 8 gives ~2^-53 relative to
// rowmax(A) colmax(B) K, i.e. fp64-level results (tests: 1e-12 vs DMMA).
//
// Pipeline (3 kernels):
//   1. ozaki_rowmax_kernel   per-row (A) / per-column (B) max |x|   (HBM bound)
//   2. ozaki_split_kernel    strided fp64 -> S packed int8 K-major slice
//                            planes [S][rows][K]; the INDEX PERMUTATION of the
//                            contraction is folded into this gather, so the
//                            GEMM kernel sees canonical operands  (HBM bound)
//   3. ozaki_gemm_kernel     128 x 64 output tile per CTA.  Warp-specialised:
//        warp 0  TMA producer: cp.async.bulk.tensor (SWIZZLE_128B) of the S
//                B-slice tiles of a k-block (double buffered) and of the A
//                slice tiles through a 4-slot ring, mbarrier full/empty;
//        warp 1  one thread issues tcgen05.mma.kind::i8 (M=128,N=64,K=32):
//                for p: for q <= S-1-p: acc[p+q] += A_p B_q; the S
//                anti-diagonal accumulators (S x 64 columns x 128 lanes of
//                int32) live in TMEM (all 512 columns); tcgen05.commit
//                releases smem slots and finally signals the epilogue;
//        warps 2-5  epilogue: tcgen05.ld the S accumulators, combine them
//                in fp64 with exact power-of-two weights, apply the row /
//                column scales and alpha/beta, scatter to the strided C
//                (output permutation folded into the store).
//
// Round 2: the same engine serves all four dtypes natively.
//   * float32 / complex64 (BASELINE configs[4]: PEPS boundary contraction):
//     S = 4 slices (6 + 7*3 = 27 bits below the row maximum, finer than the
//     24-bit significand), 10 slice products, 128 x 128 output tiles (4
//     accumulators x 128 TMEM columns); the split kernel reads the strided
//     single-precision operand directly and the epilogue rounds once to
//     float -- no widening passes through HBM.
//   * complex64 / complex128: the real GEMM of the embedding
//       A_e[m, 2k+c] = (re, im)(A[m,k]),   B_e[2k+c, 2n+d] = [[re, im], [-im, re]]
//     (C_e[m, 2n+d] is the interleaved storage of complex C); the embedding is
//     produced inside the split kernel's gather (conjugation = a sign there),
//     so complex operands cost no extra pass either.
#include <cuda.h>
#include <math.h>

#include <algorithm>

#include "internal.h"

namespace qb {

constexpr int OZ_BM = 128;     // tile rows (TMEM lanes)
constexpr int OZ_BK = 128;     // bytes (= int8 elements) of K per slice tile
constexpr int OZ_MAXS = 8;     // max slices: S * BN <= 512 TMEM columns
constexpr int OZ_ASLOTS = 4;   // A ring depth
constexpr int OZ_A_TILE = OZ_BM * OZ_BK;  // 16 KB
// tile columns per accumulator: 64 with 8 slices (double), 128 with 4 (single)

// ------------------------------------------------------------------ PTX ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map,
                                            uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
        "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
// start>>4 [0,14), LBO [16,30) = 1 (unused), SBO [32,46) = 1024>>4 (8 rows of
// 128 B), version [46,48) = 1, layout type [61,64) = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}

// --------------------------------------------------- operand preparation ----
struct SplitParams {
  const void *src;
  int32_t single;      // 1: float / complex64 source, 0: double / complex128
  int32_t role;        // 0 real; 1 complex, A-role (K doubled); 2 complex, B-role
                       // (rows and K doubled: [[re, im], [-im, re]])
  int32_t conj;        // complex operand enters conjugated
  ModeGroup rows;      // s0 = stride of the row (free) modes in src (elements)
  ModeGroup ks;        // s0 = stride of the contracted modes in src
  int64_t R, K;        // extents of the (embedded) real operand
  int64_t Rpad, Kpad;  // padded extents of the slice planes
  int32_t S;
  int32_t k_contig;    // 1: consecutive k contiguous in memory, 0: rows
  unsigned long long *rowmax;  // [Rpad] max |x| as raw double bits
  double *scale;               // [Rpad] 2^(e-6)
  int8_t *slices;              // [S][Rpad][Kpad]
};

// element (r_e, k_e) of the real operand the GEMM sees: `off` = element offset
// of the underlying (possibly complex) entry, d = r_e & 1, c = k_e & 1
__device__ __forceinline__ double oz_load(const SplitParams &P, int64_t off, int d, int c) {
  if (P.role == 0)
    return P.single ? (double)static_cast<const float *>(P.src)[off]
                    : static_cast<const double *>(P.src)[off];
  double re, im;
  if (P.single) {
    const float2 z = static_cast<const float2 *>(P.src)[off];
    re = z.x; im = z.y;
  } else {
    const double2 z = static_cast<const double2 *>(P.src)[off];
    re = z.x; im = z.y;
  }
  if (P.conj) im = -im;
  if (P.role == 1) return c ? im : re;
  // B-role: B_e[2k+c, 2n+d]
  if (c == d) return re;
  return c ? -im : im;
}

__global__ void __launch_bounds__(256)
    ozaki_rowmax_kernel(const __grid_constant__ SplitParams P) {
  __shared__ int64_t roff[64];
  __shared__ int64_t koff[256];
  __shared__ double red[8][65];
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * 64, k0 = (int64_t)blockIdx.y * 256;
  const int rsh = P.role == 2 ? 1 : 0, ksh = P.role != 0 ? 1 : 0;
  if (tid < 64) {
    int64_t r = r0 + tid, o = -1, d;
    if (r < P.R) decode2(r >> rsh, P.rows, o, d);
    roff[tid] = o;
  }
  {
    int64_t k = k0 + tid, o = -1, d;
    if (k < P.K) decode2(k >> ksh, P.ks, o, d);
    koff[tid] = o;
  }
  __syncthreads();
  if (P.k_contig) {
    const int tx = tid & 31, ty = tid >> 5;
    for (int i = 0; i < 8; ++i) {
      const int r = ty + 8 * i;
      const int64_t ro = roff[r];
      double m = 0.0;
      if (ro >= 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int64_t ko = koff[tx + 32 * j];
          if (ko >= 0) m = fmax(m, fabs(oz_load(P, ro + ko, (int)((r0 + r) & 1), (tx + 32 * j) & 1)));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (tx == 0 && ro >= 0 && m > 0.0)
        atomicMax(&P.rowmax[r0 + r], (unsigned long long)__double_as_longlong(m));
    }
  } else {
    const int tx = tid & 31, ty = tid >> 5;
    double m0 = 0.0, m1 = 0.0;
    const int64_t ro0 = roff[tx], ro1 = roff[tx + 32];
    for (int j = 0; j < 32; ++j) {
      const int64_t ko = koff[ty + 8 * j];
      if (ko >= 0) {
        const int c = (ty + 8 * j) & 1;   // k0 is even
        if (ro0 >= 0) m0 = fmax(m0, fabs(oz_load(P, ro0 + ko, tx & 1, c)));
        if (ro1 >= 0) m1 = fmax(m1, fabs(oz_load(P, ro1 + ko, tx & 1, c)));
      }
    }
    red[ty][tx] = m0; red[ty][tx + 32] = m1;
    __syncthreads();
    if (tid < 64) {
      double m = 0.0;
      for (int w = 0; w < 8; ++w) m = fmax(m, red[w][tid]);
      if (roff[tid] >= 0 && m > 0.0)
        atomicMax(&P.rowmax[r0 + tid], (unsigned long long)__double_as_longlong(m));
    }
  }
}

// tile: 32 rows x 128 k of the operand -> S x (32 x 128) int8
__global__ void __launch_bounds__(256)
    ozaki_split_kernel(const __grid_constant__ SplitParams P) {
  __shared__ double tile[32][8 * 17];  // [row][k/16][16 + 1 pad]
  __shared__ int64_t roff[32];
  __shared__ int64_t koff[128];
  __shared__ double inv_s[32];
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * 32, k0 = (int64_t)blockIdx.y * 128;
  const int rsh = P.role == 2 ? 1 : 0, ksh = P.role != 0 ? 1 : 0;
  if (tid < 32) {
    const int64_t r = r0 + tid;
    int64_t o = -1, d;
    double inv = 0.0;
    if (r < P.R) {
      decode2(r >> rsh, P.rows, o, d);
      const double mx = __longlong_as_double((long long)P.rowmax[r]);
      int e = 0;
      if (mx > 0.0) e = ilogb(mx) + 1;  // mx * 2^-e in [0.5, 1)
      inv = ldexp(1.0, 6 - e);          // x * inv in (-64, 64)
      if (blockIdx.y == 0) P.scale[r] = ldexp(1.0, e - 6);
    } else if (blockIdx.y == 0) {
      P.scale[r] = 0.0;
    }
    roff[tid] = o; inv_s[tid] = inv;
  }
  if (tid < 128) {
    const int64_t k = k0 + tid;
    int64_t o = -1, d;
    if (k < P.K) decode2(k >> ksh, P.ks, o, d);
    koff[tid] = o;
  }
  __syncthreads();
  // gather the tile, coalesced along the contiguous direction
  for (int idx = tid; idx < 32 * 128; idx += 256) {
    int r, k;
    if (P.k_contig) { k = idx & 127; r = idx >> 7; } else { r = idx & 31; k = idx >> 5; }
    const int64_t ro = roff[r], ko = koff[k];
    double x = 0.0;
    if (ro >= 0 && ko >= 0) x = oz_load(P, ro + ko, r & 1, k & 1) * inv_s[r];  // r0, k0 even
    tile[r][(k >> 4) * 17 + (k & 15)] = x;
  }
  __syncthreads();
  // each thread: one row, 16 consecutive k -> S packed 16-byte vectors
  const int seg = tid & 7, row = tid >> 3;
  double t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) t[i] = tile[row][seg * 17 + i];
  const int64_t gr = r0 + row, gk = k0 + seg * 16;
  const int64_t plane = P.Rpad * P.Kpad;
  for (int s = 0; s < P.S; ++s) {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const double q = rint(t[i]);
      t[i] = (t[i] - q) * 128.0;
      w[i >> 2] |= ((uint32_t)(uint8_t)(int8_t)(int)q) << (8 * (i & 3));
    }
    *reinterpret_cast<uint4 *>(P.slices + (int64_t)s * plane + gr * P.Kpad + gk) =
        make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ------------------------------------------------------------- the GEMM ----
struct OzGemmParams {
  ContractParams c;   // for the C scatter (mode groups m, n with s1 = stride in C)
  int32_t S;
  int32_t nkb;        // k-blocks of 128
  int32_t cplx;       // C is complex (interleaved): column 2n+d of the real GEMM
  int64_t Me, Ne;     // extents of the real GEMM
  const double *scaleA;   // [Mpad]
  const double *scaleB;   // [Npad]
};

template <int BN>
struct OzSmem {
  static constexpr int B_TILE = BN * OZ_BK;               // 8 / 16 KB
  static constexpr int NS = 512 / BN;                     // slices that fit TMEM
  static constexpr size_t TILES = (size_t)OZ_ASLOTS * OZ_A_TILE + 2 * (size_t)NS * B_TILE;
  static constexpr size_t BYTES = TILES + 1024 /*align*/ + 8192 /*tables, barriers*/;
};

template <int BN, typename OutT>
__global__ void __launch_bounds__(192, 1)
    ozaki_gemm_kernel(const __grid_constant__ CUtensorMap mapA,
                      const __grid_constant__ CUtensorMap mapB,
                      const __grid_constant__ OzGemmParams P) {
  constexpr int B_TILE = OzSmem<BN>::B_TILE;
  constexpr int NS = OzSmem<BN>::NS;
  extern __shared__ unsigned char oz_smem_raw[];
  unsigned char *base = reinterpret_cast<unsigned char *>(
      ((uintptr_t)oz_smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char *sA = base;                                   // 4 x 16 KB
  unsigned char *sB = base + (size_t)OZ_ASLOTS * OZ_A_TILE;   // 2 x NS x B_TILE
  unsigned char *aux = base + OzSmem<BN>::TILES;
  uint64_t *fullA = reinterpret_cast<uint64_t *>(aux);        // [4]
  uint64_t *emptyA = fullA + OZ_ASLOTS;                       // [4]
  uint64_t *fullB = emptyA + OZ_ASLOTS;                       // [2]
  uint64_t *emptyB = fullB + 2;                               // [2]
  uint64_t *tmem_full = emptyB + 2;                           // [1]
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(tmem_full + 1);
  int64_t *offCm = reinterpret_cast<int64_t *>(aux + 256);    // [128]
  int64_t *offCn = offCm + OZ_BM;                             // [BN]
  double *sclM = reinterpret_cast<double *>(offCn + BN);      // [128]
  double *sclN = sclM + OZ_BM;                                // [BN]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = P.S, nkb = P.nkb;
  const int tm = blockIdx.x, tn = blockIdx.y;

  if (tid == 0) {
    for (int i = 0; i < OZ_ASLOTS; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_base_s)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // epilogue tables (warps 2..5 = 128 threads); offsets in units of OutT
  if (warp >= 2) {
    const int e = tid - 64;
    const int csh = P.cplx ? 1 : 0;
    {
      const int64_t m = (int64_t)tm * OZ_BM + e;
      int64_t oa, oc = -1;
      if (m < P.Me) { decode2(m, P.c.m, oa, oc); oc <<= csh; }
      offCm[e] = oc;
      sclM[e] = P.scaleA[(int64_t)tm * OZ_BM + e];
    }
    if (e < BN) {
      const int64_t ne = (int64_t)tn * BN + e;
      int64_t ob, oc = -1;
      if (ne < P.Ne) { decode2(ne >> csh, P.c.n, ob, oc); oc = (oc << csh) + (csh ? (ne & 1) : 0); }
      offCn[e] = oc;
      sclN[e] = P.scaleB[(int64_t)tn * BN + e];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;

  if (warp == 0) {
    // =========================== TMA producer ==============================
    if (lane == 0) {
      int ia = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int sb = kb & 1;
        mbar_wait(&emptyB[sb], ((kb >> 1) & 1) ^ 1);
        mbar_expect_tx(&fullB[sb], (uint32_t)S * B_TILE);
        for (int q = 0; q < S; ++q)
          tma_load_3d(sB + ((size_t)sb * NS + q) * B_TILE, &mapB, &fullB[sb],
                      kb * OZ_BK, tn * BN, q);
        for (int p = 0; p < S; ++p, ++ia) {
          const int sa = ia % OZ_ASLOTS;
          mbar_wait(&emptyA[sa], ((ia / OZ_ASLOTS) & 1) ^ 1);
          mbar_expect_tx(&fullA[sa], OZ_A_TILE);
          tma_load_3d(sA + (size_t)sa * OZ_A_TILE, &mapA, &fullA[sa], kb * OZ_BK,
                      tm * OZ_BM, p);
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ================================
    if (lane == 0) {
      // kind::i8: D = S32 (2<<4), A = INT8 (1<<7), B = INT8 (1<<10), K-major
      // both, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
      int ia = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int sb = kb & 1;
        mbar_wait(&fullB[sb], (kb >> 1) & 1);
        tc_fence_after();
        for (int p = 0; p < S; ++p, ++ia) {
          const int sa = ia % OZ_ASLOTS;
          mbar_wait(&fullA[sa], (ia / OZ_ASLOTS) & 1);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + (size_t)sa * OZ_A_TILE));
          for (int q = 0; q + p < S; ++q) {
            const uint64_t bdesc =
                umma_desc_sw128(smem_u32(sB + ((size_t)sb * NS + q) * B_TILE));
            const uint32_t dcol = tmem_base + (uint32_t)(p + q) * BN;
#pragma unroll
            for (int k4 = 0; k4 < OZ_BK / 32; ++k4) {
              // first touch of accumulator d = p+q is (kb=0, p=0, k4=0)
              const uint32_t accum = (kb > 0 || p > 0 || k4 > 0) ? 1u : 0u;
              umma_i8(dcol, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), idesc,
                      accum);
            }
          }
          umma_commit(&emptyA[sa]);  // slot free once these MMAs retire
        }
        umma_commit(&emptyB[sb]);
      }
      umma_commit(tmem_full);
    }
  } else {
    // ============================== epilogue ================================
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;   // row of the tile = TMEM lane
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int64_t om = offCm[row];
    const double sm = sclM[row] * P.c.alpha;
    OutT *C = static_cast<OutT *>(P.c.C);
    const double beta = P.c.beta;
    for (int cc = 0; cc < BN; cc += 16) {
      double acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.0;
      double w = ldexp(1.0, -7 * (S - 1));
      for (int d = S - 1; d >= 0; --d) {
        int32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(d * BN + cc), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += w * (double)v[i];
        w *= 128.0;
      }
      if (om >= 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int64_t on = offCn[cc + i];
          if (on >= 0) {
            double val = acc[i] * sm * sclN[cc + i];
            if (beta != 0.0) val += beta * (double)C[om + on];
            C[om + on] = (OutT)val;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(512));
  }
}

// ------------------------------------------------------------------ host ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static int make_slice_map(CUtensorMap *map, const int8_t *ptr, int64_t Kpad, int64_t Rpad,
                          int S, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable");
    return 2000;
  }
  cuuint64_t dims[3] = {(cuuint64_t)Kpad, (cuuint64_t)Rpad, (cuuint64_t)S};
  cuuint64_t strides[2] = {(cuuint64_t)Kpad, (cuuint64_t)(Kpad * Rpad)};
  cuuint32_t box[3] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)ptr, dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d", (int)r);
    return 2000 + (int)r;
  }
  return 0;
}

struct OzGeom {
  int S, BN, cplx, single;
  int64_t Me, Ne, Ke;            // extents of the real GEMM
  int64_t Mpad, Npad, Kpad;
  int64_t off_rmA, off_rmB, off_scA, off_scB, off_slA, off_slB, total;  // bytes
};

static int oz_slices(bool single) {
  static int s8 = [] {
    const char *e = getenv("QB_OZAKI_SLICES");
    int v = e ? atoi(e) : 8;
    return std::max(2, std::min(v, OZ_MAXS));
  }();
  static int s4 = [] {
    const char *e = getenv("QB_OZAKI_SLICES_SINGLE");
    int v = e ? atoi(e) : 4;
    return std::max(2, std::min(v, 4));
  }();
  return single ? s4 : s8;
}

static void oz_geometry(const PairPlan &plan, OzGeom &g) {
  const ContractParams &p = plan.p;
  auto up = [](int64_t x, int64_t a) { return (x + a - 1) / a * a; };
  g.single = (plan.dtype == QB_F32 || plan.dtype == QB_C64) ? 1 : 0;
  g.cplx = dtype_is_complex(plan.dtype) ? 1 : 0;
  g.S = oz_slices(g.single);
  g.BN = g.single ? 128 : 64;
  g.Me = p.M;
  g.Ne = g.cplx ? 2 * p.N : p.N;
  g.Ke = g.cplx ? 2 * p.K : p.K;
  g.Mpad = up(g.Me, OZ_BM);
  g.Npad = up(g.Ne, OZ_BM);  // multiple of 128 so the split kernel tiles evenly
  g.Kpad = up(g.Ke, OZ_BK);
  int64_t off = 0;
  g.off_rmA = off; off += up(g.Mpad * 8, 1024);
  g.off_rmB = off; off += up(g.Npad * 8, 1024);
  g.off_scA = off; off += up(g.Mpad * 8, 1024);
  g.off_scB = off; off += up(g.Npad * 8, 1024);
  g.off_slA = off; off += up((int64_t)g.S * g.Mpad * g.Kpad, 1024);
  g.off_slB = off; off += up((int64_t)g.S * g.Npad * g.Kpad, 1024);
  g.total = off;
}

bool ozaki_eligible(const PairPlan &plan) {
  const ContractParams &p = plan.p;
  const int64_t f = dtype_is_complex(plan.dtype) ? 2 : 1;
  return p.nbatch == 1 && !p.dA && p.M >= 128 && f * p.N >= 64 && f * p.K >= 128 &&
         f * p.K < 65536 && !plan.empty_out && !plan.zero_fill;
}

int64_t ozaki_workspace_bytes(const PairPlan &plan) {
  OzGeom g;
  oz_geometry(plan, g);
  return g.total;
}

template <int BN, typename OutT>
static int oz_launch_gemm(const CUtensorMap &mapA, const CUtensorMap &mapB,
                          const OzGemmParams &gp, dim3 grid, cudaStream_t st) {
  auto kern = ozaki_gemm_kernel<BN, OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    QB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)OzSmem<BN>::BYTES));
    attr_set = true;
  }
  kern<<<grid, 192, OzSmem<BN>::BYTES, st>>>(mapA, mapB, gp);
  QB_LAUNCH_CHECK();
  return 0;
}

int launch_contract_ozaki(const PairPlan &plan, void *workspace, cudaStream_t st) {
  const ContractParams &p = plan.p;
  OzGeom g;
  oz_geometry(plan, g);
  unsigned char *ws = static_cast<unsigned char *>(workspace);
  if (((uintptr_t)ws & 255) != 0) {
    set_error("ozaki workspace must be 256-byte aligned");
    return -10;
  }
  QB_CUDA_CHECK(cudaMemsetAsync(ws, 0, g.off_scA, st));  // row/col maxima = 0
  auto prep = [&](const void *src, int role, int conj, const ModeGroup &rows,
                  const ModeGroup &ks, bool k_second, int64_t R, int64_t Rpad,
                  int64_t off_rm, int64_t off_sc, int64_t off_sl) -> int {
    SplitParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.src = src;
    sp.single = g.single;
    sp.role = role;
    sp.conj = conj;
    sp.rows.n = rows.n; sp.ks.n = ks.n;
    for (int i = 0; i < rows.n; ++i) {
      sp.rows.ext[i] = rows.ext[i];
      sp.rows.s0[i] = rows.s0[i];
    }
    for (int i = 0; i < ks.n; ++i) {
      sp.ks.ext[i] = ks.ext[i];
      sp.ks.s0[i] = k_second ? ks.s1[i] : ks.s0[i];
    }
    sp.R = R; sp.K = g.Ke; sp.Rpad = Rpad; sp.Kpad = g.Kpad; sp.S = g.S;
    const int64_t ksm = ks.n ? sp.ks.s0[0] : 1, rsm = rows.n ? sp.rows.s0[0] : (int64_t)1 << 60;
    sp.k_contig = (ksm <= rsm) ? 1 : 0;
    sp.rowmax = reinterpret_cast<unsigned long long *>(ws + off_rm);
    sp.scale = reinterpret_cast<double *>(ws + off_sc);
    sp.slices = reinterpret_cast<int8_t *>(ws + off_sl);
    dim3 g1((unsigned)((R + 63) / 64), (unsigned)((g.Ke + 255) / 256));
    ozaki_rowmax_kernel<<<g1, 256, 0, st>>>(sp);
    QB_LAUNCH_CHECK();
    dim3 g2((unsigned)(Rpad / 32), (unsigned)(g.Kpad / 128));
    ozaki_split_kernel<<<g2, 256, 0, st>>>(sp);
    QB_LAUNCH_CHECK();
    return 0;
  };
  int rc;
  // A: rows = m modes (stride in A = s0), k modes (stride in A = s0)
  if ((rc = prep(p.A, g.cplx ? 1 : 0, p.conjA, p.m, p.k, false, g.Me, g.Mpad,
                 g.off_rmA, g.off_scA, g.off_slA)))
    return rc;
  // B: rows = n modes (stride in B = s0), k modes (stride in B = s1)
  if ((rc = prep(p.B, g.cplx ? 2 : 0, p.conjB, p.n, p.k, true, g.Ne, g.Npad,
                 g.off_rmB, g.off_scB, g.off_slB)))
    return rc;

  CUtensorMap mapA, mapB;
  if ((rc = make_slice_map(&mapA, reinterpret_cast<int8_t *>(ws + g.off_slA), g.Kpad, g.Mpad,
                           g.S, OZ_BM)))
    return rc;
  if ((rc = make_slice_map(&mapB, reinterpret_cast<int8_t *>(ws + g.off_slB), g.Kpad, g.Npad,
                           g.S, g.BN)))
    return rc;
  OzGemmParams gp;
  gp.c = p;
  gp.S = g.S;
  gp.nkb = (int)(g.Kpad / OZ_BK);
  gp.cplx = g.cplx;
  gp.Me = g.Me; gp.Ne = g.Ne;
  gp.scaleA = reinterpret_cast<const double *>(ws + g.off_scA);
  gp.scaleB = reinterpret_cast<const double *>(ws + g.off_scB);
  dim3 grid((unsigned)(g.Mpad / OZ_BM), (unsigned)((g.Ne + g.BN - 1) / g.BN));
  if (g.single) return oz_launch_gemm<128, float>(mapA, mapB, gp, grid, st);
  return oz_launch_gemm<64, double>(mapA, mapB, gp, grid, st);
}

}  // namespace qb
