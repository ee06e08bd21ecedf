// Pairwise tensor contraction on the fp64 tensor cores (DMMA) of sm_100a.
//
//   C[m, n, batch] = sum_k op(A)[m, k, batch] * op(B)[k, n, batch]
//
// where m, n, k and batch are *groups of tensor modes* with arbitrary
// strides (see plan.h).  No operand is ever transposed or copied in global
// memory: a CTA turns its tile's logical indices into element offsets through
// small per-tile tables held in shared memory, and the cp.async tile loads
// gather straight from the strided tensors, so the index permutation that
// numpy.tensordot performs as a separate transpose-copy
// (quimb -> cotengra -> autoray do("tensordot")) is folded into the load.
//
// Pipeline: STAGES-deep cp.async ring (16-byte copies whenever the planner
// proved pairs of elements contiguous, 8-byte otherwise), one __syncthreads
// per k-block, warp tiles of m16n8k8 DMMA, accumulators in registers,
// direct (optionally 16-byte) strided stores.  Split-K writes partial tiles
// to a workspace that a second kernel reduces deterministically.
//
// complex128 runs through the same kernel as a real GEMM of twice the N and
// K extents:  A is read as reals with (re,im) as the fastest contracted
// mode, C is written as reals with (re,im) as the fastest n mode, and the
// B tile (kept as complex in shared memory, half the bytes) is expanded on
// the fly in the fragment read:
//     Bhat[(k,a),(n,c)] = sign(a,c) * B[k,n].component(a xor c)
// sign = -1 for (a=1,c=0); conjugation of either operand only changes the
// per-thread sign rule, so it is free.
#include "common.cuh"
#include "plan.h"

namespace qb {

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES>
struct KernelCfg {
  static constexpr int NT = WARPS_M * WARPS_N * 32;
  static constexpr int WM = BM / WARPS_M;
  static constexpr int WN = BN / WARPS_N;
  static constexpr int MT = WM / 16;
  static constexpr int NT8 = WN / 8;
  // tile storage: either [row][k] with pitch BK+4 or [k][row] with pitch
  // rows+4; both are bank-conflict free for the DMMA fragment reads
  // (pitch == 4 mod 16 doubles).  The complex B tile is [k/2][n/2] complex
  // with pitch BN/2+2 complex, which needs fewer doubles than either.
  static constexpr int A_ELEMS =
      (BM * (BK + 4) > BK * (BM + 4)) ? BM * (BK + 4) : BK * (BM + 4);
  static constexpr int B_ELEMS =
      (BN * (BK + 4) > BK * (BN + 4)) ? BN * (BK + 4) : BK * (BN + 4);
  static constexpr int BC_PITCH = BN / 2 + 2;  // complex elements
  static constexpr size_t SMEM =
      (size_t)STAGES * (A_ELEMS + B_ELEMS) * 8       // tiles
      + (size_t)(2 * BM + 2 * BN) * 8                // m/n offset tables
      + (size_t)(2 * 2 * BK) * 8;                    // k offset tables (x2)
};

// gather one ROWS x BK tile of 8-byte elements into shared memory
template <int ROWS, int BK, int NT, int KM = -1>
__device__ __forceinline__ void load_tile(
    double *__restrict__ sm, const double *__restrict__ gbase,
    const int64_t *__restrict__ row_off, const int64_t *__restrict__ k_off,
    int vec, int thr, int tid) {
  // vec: 0 scalar, 1 pairs along rows, 2 pairs along k.  thr: 1 walk k.
  // smem layout: vec==1 or (vec==0 && thr==0) -> [k][row]; else [row][k]
  // KM >= 0: the layout is a compile-time constant (specialised kernels)
  const bool kmajor = KM >= 0 ? (KM == 1) : ((vec == 2) || (vec == 0 && thr == 1));
  const int s_row = kmajor ? (BK + 4) : 1;
  const int s_k = kmajor ? 1 : (ROWS + 4);
  if (vec == 0) {
    constexpr int TOTAL = ROWS * BK;
#pragma unroll
    for (int c0 = 0; c0 < TOTAL; c0 += NT) {
      int c = c0 + tid;
      if (TOTAL % NT != 0 && c >= TOTAL) break;
      int r, k;
      if (thr) { k = c % BK; r = c / BK; } else { r = c % ROWS; k = c / ROWS; }
      int64_t ro = row_off[r], ko = k_off[k];
      bool ok = (ro >= 0) && (ko >= 0);
      const double *src = ok ? (gbase + ro + ko) : gbase;
      cp_async8(smem_u32(sm + r * s_row + k * s_k), src, ok ? 8 : 0);
    }
  } else if (vec == 2) {
    constexpr int TOTAL = ROWS * BK / 2;
    constexpr int KC = BK / 2;
#pragma unroll
    for (int c0 = 0; c0 < TOTAL; c0 += NT) {
      int c = c0 + tid;
      if (TOTAL % NT != 0 && c >= TOTAL) break;
      int r, kc;
      if (thr) { kc = c % KC; r = c / KC; } else { r = c % ROWS; kc = c / ROWS; }
      int k = kc * 2;
      int64_t ro = row_off[r], ko = k_off[k];
      bool ok = (ro >= 0) && (ko >= 0);
      const double *src = ok ? (gbase + ro + ko) : gbase;
      cp_async16(smem_u32(sm + r * s_row + k * s_k), src, ok ? 16 : 0);
    }
  } else {
    constexpr int TOTAL = ROWS * BK / 2;
    constexpr int RC = ROWS / 2;
#pragma unroll
    for (int c0 = 0; c0 < TOTAL; c0 += NT) {
      int c = c0 + tid;
      if (TOTAL % NT != 0 && c >= TOTAL) break;
      int rc, k;
      if (thr) { k = c % BK; rc = c / BK; } else { rc = c % RC; k = c / RC; }
      int r = rc * 2;
      int64_t ro = row_off[r], ko = k_off[k];
      bool ok = (ro >= 0) && (ko >= 0);
      const double *src = ok ? (gbase + ro + ko) : gbase;
      cp_async16(smem_u32(sm + r * s_row + k * s_k), src, ok ? 16 : 0);
    }
  }
}

// gather the complex B tile: (BK/2) x (BN/2) complex elements, layout
// [k][n] with pitch PITCH complex; offsets are in doubles (already x2)
template <int NC, int KC, int PITCH, int NT>
__device__ __forceinline__ void load_tile_cplx(
    double *__restrict__ sm, const double *__restrict__ gbase,
    const int64_t *__restrict__ n_off, const int64_t *__restrict__ k_off,
    int thr, int tid) {
  constexpr int TOTAL = NC * KC;
#pragma unroll
  for (int c0 = 0; c0 < TOTAL; c0 += NT) {
    int c = c0 + tid;
    if (TOTAL % NT != 0 && c >= TOTAL) break;
    int n, k;
    if (thr) { k = c % KC; n = c / KC; } else { n = c % NC; k = c / NC; }
    int64_t no = n_off[n], ko = k_off[k];
    bool ok = (no >= 0) && (ko >= 0);
    const double *src = ok ? (gbase + no + ko) : gbase;
    cp_async16(smem_u32(sm + (k * PITCH + n) * 2), src, ok ? 16 : 0);
  }
}

__device__ __forceinline__ double flip_sign(double v, unsigned mask_hi) {
  // mask_hi is 0 or 0x80000000
  return __hiloint2double(__double2hiint(v) ^ (int)mask_hi, __double2loint(v));
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// tuning: stamp phase `ph` of segment `seg` of this CTA (8 phases x 4 segments)
#define QB_TRACE(ph)                                                        \
  do {                                                                      \
    if (p.trace && threadIdx.x == 0 && sk.seg < 4)                          \
      p.trace[((size_t)sk.bid * 4 + sk.seg) * 8 + (ph)] = global_ns();      \
  } while (0)

// stream-K bookkeeping of one tile segment
struct SkSeg {
  int mode;        // 0: classic tile, 1: tail (write partial), 2: head (sum peers)
  int bid;         // this CTA (partial slot)
  int peer0, peer1;  // head: CTAs [peer0, peer1) hold the rest of the tile
  double *part;    // [grid][BM*BN] partial accumulators
  int *flags;      // [grid]
  int seg;         // ordinal of this segment within the CTA (trace only)
};

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES,
          bool CPLX, int LA = -1, int LB = -1, int DBG = 0>
__device__ __forceinline__ void contract_tile(const ContractParams &p, int pid,
                                              int64_t zb, int ks, int64_t kbeg,
                                              int64_t kend, const SkSeg sk) {
  using Cfg = KernelCfg<BM, BN, BK, WARPS_M, WARPS_N, STAGES>;
  constexpr int NT = Cfg::NT;
  constexpr int MT = Cfg::MT, NT8 = Cfg::NT8;
  constexpr int NCB = BN / 2, KCB = BK / 2;  // complex B tile extents

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sA = reinterpret_cast<double *>(smem_raw);
  double *sB = sA + (size_t)STAGES * Cfg::A_ELEMS;
  int64_t *offAm = reinterpret_cast<int64_t *>(sB + (size_t)STAGES * Cfg::B_ELEMS);
  int64_t *offCm = offAm + BM;
  int64_t *offBn = offCm + BM;
  int64_t *offCn = offBn + BN;
  int64_t *ktabA = offCn + BN;      // [2][BK]
  int64_t *ktabB = ktabA + 2 * BK;  // [2][BK]

  QB_TRACE(0);
  // let the next kernel of the stream become resident as SMs free up ...
  asm volatile("griddepcontrol.launch_dependents;");
  // ... and do not touch global memory before the previous one has finished
  // (pointer-array batches read their arrays right away)
  if (p.dA) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / WARPS_N) * Cfg::WM;
  const int wn0 = (warp % WARPS_N) * Cfg::WN;

  // effective real extents
  const int64_t Nh = CPLX ? 2 * p.N : p.N;
  const int64_t Kh = CPLX ? 2 * p.K : p.K;

  // ---- which tile / batch / k-split am I -------------------------------
  int tm, tn;
  {
    // grouped rasterisation: walk 8 m-tiles down before moving along n so
    // that concurrently resident CTAs share B panels in L2
    const int G = 8;
    int width = G * p.tiles_n;
    int group = pid / width;
    int first_m = group * G;
    int gsz = min(p.tiles_m - first_m, G);
    int rem = pid - group * width;
    tm = first_m + rem % gsz;
    tn = rem / gsz;
  }
  const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;

  const double *A = static_cast<const double *>(p.A);
  const double *B = static_cast<const double *>(p.B);
  double *C = static_cast<double *>(p.C);
  constexpr int ES = CPLX ? 2 : 1;  // doubles per element
  if (p.dA) {
    A = static_cast<const double *>(p.dA[zb]);
    B = static_cast<const double *>(p.dB[zb]);
    C = static_cast<double *>(p.dC[zb]);
  } else if (p.b.n) {
    int64_t oa, ob, oc = 0;
    decode2(zb, p.b, oa, ob);
    {  // C batch offset
      int64_t r = zb;
      for (int i = 0; i < p.b.n; ++i) {
        int64_t e = p.b.ext[i];
        int64_t q = r / e;
        oc += (r - q * e) * p.bsC[i];
        r = q;
      }
    }
    A += oa * ES; B += ob * ES; C += oc * ES;
  }

  // ---- per-tile offset tables (in doubles) -----------------------------
  for (int i = tid; i < BM; i += NT) {
    int64_t m = m0 + i, oa = -1, oc = -1;
    if (m < p.M) { decode2(m, p.m, oa, oc); oa *= ES; oc *= ES; }
    offAm[i] = oa; offCm[i] = oc;
  }
  for (int i = tid; i < BN; i += NT) {
    int64_t nh = n0 + i, ob = -1, oc = -1;
    if (nh < Nh) {
      if (CPLX) {
        decode2(nh >> 1, p.n, ob, oc);
        ob *= 2; oc = oc * 2 + (nh & 1);
      } else {
        decode2(nh, p.n, ob, oc);
      }
    }
    // complex: offBn is indexed by the complex column i/2 (even i writes it)
    if (CPLX) { if ((i & 1) == 0) offBn[i >> 1] = ob; }
    else offBn[i] = ob;
    offCn[i] = oc;
  }
  const int nkb = (int)((kend - kbeg + BK - 1) / BK);

  auto fill_ktab = [&](int kb) {
    // offsets of k-block kb into parity slot kb&1 (threads < BK)
    if (tid < BK) {
      int64_t kh = kbeg + (int64_t)kb * BK + tid, oa = -1, ob = -1;
      if (kb < nkb && kh < kend) {
        if (CPLX) {
          decode2(kh >> 1, p.k, oa, ob);
          oa = oa * 2 + (kh & 1); ob *= 2;
        } else {
          decode2(kh, p.k, oa, ob);
        }
      }
      ktabA[(kb & 1) * BK + tid] = oa;
      if (CPLX) { if ((tid & 1) == 0) ktabB[(kb & 1) * BK + (tid >> 1)] = ob; }
      else ktabB[(kb & 1) * BK + tid] = ob;
    }
  };
  auto issue = [&](int kb) {
    if (kb < nkb) {
      const int st = kb % STAGES;
      load_tile<BM, BK, NT, LA>(sA + (size_t)st * Cfg::A_ELEMS, A, offAm,
                                ktabA + (kb & 1) * BK, CPLX ? 2 : p.vecA, p.thrA, tid);
      if (CPLX)
        load_tile_cplx<NCB, KCB, Cfg::BC_PITCH, NT>(
            sB + (size_t)st * Cfg::B_ELEMS, B, offBn, ktabB + (kb & 1) * BK,
            p.thrB, tid);
      else
        load_tile<BN, BK, NT, LB>(sB + (size_t)st * Cfg::B_ELEMS, B, offBn,
                                  ktabB + (kb & 1) * BK, p.vecB, p.thrB, tid);
    }
    cp_async_commit();
  };

  // prologue: tables for block 0, then STAGES-1 loads in flight
  fill_ktab(0);
  if (!p.dA) asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  QB_TRACE(1);
#pragma unroll 1
  for (int s = 0; s < STAGES - 1; ++s) {
    issue(s);
    fill_ktab(s + 1);
    __syncthreads();
  }

  QB_TRACE(2);
  double acc[MT][NT8][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT8; ++j)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[i][j][v] = 0.0;

  const int vecA = CPLX ? 2 : p.vecA;
  const bool a_kmajor = LA >= 0 ? (LA == 1) : ((vecA == 2) || (vecA == 0 && p.thrA == 1));
  const bool b_kmajor = LB >= 0 ? (LB == 1) : ((p.vecB == 2) || (p.vecB == 0 && p.thrB == 1));
  const int sAm = a_kmajor ? (BK + 4) : 1, sAk = a_kmajor ? 1 : (BM + 4);
  const int sBn = b_kmajor ? (BK + 4) : 1, sBk = b_kmajor ? 1 : (BN + 4);
  // complex B expansion: per-thread component and sign (a = t&1, c = g&1)
  const int ca = t & 1, cc = g & 1, comp = ca ^ cc;
  unsigned sgn_mask = 0;
  if (CPLX) {
    int neg = (ca == 1 && cc == 0) ? 1 : 0;
    if (p.conjB && comp == 1) neg ^= 1;
    if (p.conjA && ca == 1) neg ^= 1;
    sgn_mask = neg ? 0x80000000u : 0u;
  }

  const bool early = ((warp >> 2) & 1) == 0 || (NT <= 128);
  // DBG (tuning builds only, QB_DBG): 1 no tile loads, 2 no barriers,
  // 4 fragments read once -- results are garbage, only the timing is used
  constexpr int dbg = DBG;
  double af[MT][4], bf[NT8][2];
#pragma unroll 1
  for (int kb = 0; kb < nkb; ++kb) {
    if (!(dbg & 2)) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();  // stage kb landed; stage kb-1 free; next ktab visible
    }
    if (kb == 0) QB_TRACE(3);
    // the two warps that share a scheduler (warp, warp+4) issue their
    // gathers at different points of the k-block so that one of them always
    // has DMMAs ready while the other does address arithmetic
    if (early && !(dbg & 1)) issue(kb + STAGES - 1);
    fill_ktab(kb + STAGES);
    const double *tA = sA + (size_t)(kb % STAGES) * Cfg::A_ELEMS;
    const double *tB = sB + (size_t)(kb % STAGES) * Cfg::B_ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 8) {
      if (kk == 8 && !early && !(dbg & 1)) issue(kb + STAGES - 1);
      if (!(dbg & 4) || kb == 0) {
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int r = wm0 + i * 16 + g;
        af[i][0] = tA[r * sAm + (kk + t) * sAk];
        af[i][1] = tA[(r + 8) * sAm + (kk + t) * sAk];
        af[i][2] = tA[r * sAm + (kk + t + 4) * sAk];
        af[i][3] = tA[(r + 8) * sAm + (kk + t + 4) * sAk];
      }
#pragma unroll
      for (int j = 0; j < NT8; ++j) {
        const int c = wn0 + j * 8 + g;
        if (CPLX) {
          const int kc0 = (kk + t) >> 1, kc1 = (kk + t + 4) >> 1, nc = c >> 1;
          bf[j][0] = flip_sign(tB[(kc0 * Cfg::BC_PITCH + nc) * 2 + comp], sgn_mask);
          bf[j][1] = flip_sign(tB[(kc1 * Cfg::BC_PITCH + nc) * 2 + comp], sgn_mask);
        } else {
          bf[j][0] = tB[c * sBn + (kk + t) * sBk];
          bf[j][1] = tB[c * sBn + (kk + t + 4) * sBk];
        }
      }
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT8; ++j) dmma_16x8x8(acc[i][j], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  QB_TRACE(4);

  // ---- stream-K: tail segments park their accumulators, heads collect ----
  if (sk.mode == 1) {
    double *dst = sk.part + (size_t)sk.bid * (BM * BN);
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT8; ++j)
#pragma unroll
        for (int v = 0; v < 4; ++v)
          __stcg(&dst[((i * NT8 + j) * 4 + v) * NT + tid], acc[i][j][v]);
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(&sk.flags[sk.bid], 1);
    QB_TRACE(6);
    return;
  }
  if (sk.mode == 2) {
    for (int c = sk.peer0; c < sk.peer1; ++c) {
      if (tid == 0) {
        while (atomicAdd(&sk.flags[c], 0) == 0) __nanosleep(100);
        atomicExch(&sk.flags[c], 0);  // left clean for the next launch
      }
      __syncthreads();
      __threadfence();
      const double *src = sk.part + (size_t)c * (BM * BN);
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT8; ++j)
#pragma unroll
          for (int v = 0; v < 4; ++v)
            acc[i][j][v] += __ldcg(&src[((i * NT8 + j) * 4 + v) * NT + tid]);
    }
  }

  QB_TRACE(5);
  // ---- epilogue ----------------------------------------------------------
  if (p.splitk > 1) {
    // canonical [split][batch][M][Nh] partial buffer
    double *P = p.partial + ((int64_t)ks * p.nbatch + zb) * p.M * Nh;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT8; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int64_t m = m0 + wm0 + i * 16 + g + h * 8;
          int64_t n = n0 + wn0 + j * 8 + 2 * t;
          if (m < p.M) {
            if (n < Nh) P[m * Nh + n] = acc[i][j][2 * h];
            if (n + 1 < Nh) P[m * Nh + n + 1] = acc[i][j][2 * h + 1];
          }
        }
    return;
  }
  const bool vecC = CPLX ? true : (p.vecC != 0);
  const double alpha = p.alpha, beta = p.beta;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = wm0 + i * 16 + g + h * 8;
      const int64_t om = offCm[r];
      if (om < 0) continue;
#pragma unroll
      for (int j = 0; j < NT8; ++j) {
        const int c = wn0 + j * 8 + 2 * t;
        const int64_t on0 = offCn[c], on1 = offCn[c + 1];
        double v0 = alpha * acc[i][j][2 * h], v1 = alpha * acc[i][j][2 * h + 1];
        if (vecC) {
          // (c, c+1) are contiguous and 16B aligned (planner / complex pair)
          if (on0 >= 0) {
            double2 *dst = reinterpret_cast<double2 *>(C + om + on0);
            if (beta != 0.0) {
              double2 old = *dst;
              v0 += beta * old.x; v1 += beta * old.y;
            }
            *dst = make_double2(v0, v1);
          }
        } else {
          if (on0 >= 0) {
            if (beta != 0.0) v0 += beta * C[om + on0];
            C[om + on0] = v0;
          }
          if (on1 >= 0) {
            if (beta != 0.0) v1 += beta * C[om + on1];
            C[om + on1] = v1;
          }
        }
      }
    }
  QB_TRACE(6);
}

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES,
          bool CPLX, int LA = -1, int LB = -1, int DBG = 0>
__global__ void __launch_bounds__(WARPS_M *WARPS_N * 32)
    contract_f64_kernel(const __grid_constant__ ContractParams p) {
  const int64_t Kh = CPLX ? 2 * p.K : p.K;
  const int ks = blockIdx.z;
  const int64_t kbeg = (int64_t)ks * p.k_per_split;  // in real-k units
  const int64_t kend = min(Kh, kbeg + p.k_per_split);
  SkSeg sk{0, (int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)), 0, 0, nullptr, nullptr, 0};
  contract_tile<BM, BN, BK, WARPS_M, WARPS_N, STAGES, CPLX, LA, LB, DBG>(
      p, blockIdx.x, blockIdx.y, ks, kbeg, kend, sk);
}

// Stream-K: one persistent CTA per SM; the (tile, k-block) iteration space is
// cut into gridDim.x equal contiguous ranges, so every SM gets the same number
// of k-blocks whatever the tile count (no wave quantisation).  A range that
// starts inside a tile is a TAIL: its accumulators go to a per-CTA slot of the
// workspace; the CTA whose range holds the first k-block of that tile (the
// HEAD, processed last in its range, i.e. long after the tails were written)
// adds them in a fixed order and runs the normal epilogue.
template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES,
          bool CPLX, int LA = -1, int LB = -1>
__global__ void __launch_bounds__(WARPS_M *WARPS_N * 32)
    contract_f64_streamk_kernel(const __grid_constant__ ContractParams p) {
  const int64_t Kh = CPLX ? 2 * p.K : p.K;
  const int64_t nkbT = (Kh + BK - 1) / BK;
  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
  const int G = gridDim.x, bid = blockIdx.x;
  // equal shares of COST: every tile carries sk_head extra k-blocks (fix-up
  // + epilogue) that are charged to whoever runs its first k-block
  const int64_t H = p.sk_head, span = nkbT + H, total_cost = tiles * span;
  auto start = [&](int c) -> int64_t {
    const int64_t x = total_cost * c / G, t = x / span;
    return t * nkbT + max(x - t * span - H, (int64_t)0);
  };
  const int64_t u0 = start(bid), u1 = start(bid + 1);
  double *part = p.partial;
  int *flags = p.flags;
  int seg = 0;
  for (int64_t u = u0; u < u1;) {
    const int64_t tile = u / nkbT, kb0 = u - tile * nkbT;
    const int64_t kb1 = min(nkbT, kb0 + (u1 - u));
    SkSeg sk{0, bid, 0, 0, part, flags, seg++};
    if (kb0 > 0) sk.mode = 1;
    else if (kb1 < nkbT) {
      // peers: CTAs after this one whose range starts inside this tile
      sk.mode = 2;
      sk.peer0 = bid + 1;
      int c = bid + 1;
      while (c < G && start(c) < (tile + 1) * nkbT) ++c;
      sk.peer1 = c;
    }
    contract_tile<BM, BN, BK, WARPS_M, WARPS_N, STAGES, CPLX, LA, LB>(
        p, (int)tile, 0, 0, kb0 * BK, min(Kh, kb1 * BK), sk);
    u += kb1 - kb0;
    __syncthreads();  // shared memory is reused by the next segment
  }
}

// deterministic split-K reduction + strided scatter into C
template <bool CPLX>
__global__ void splitk_reduce_f64_kernel(const __grid_constant__ ContractParams p) {
  const int64_t Nh = CPLX ? 2 * p.N : p.N;
  const int64_t MN = p.M * Nh;
  const int64_t total = MN * p.nbatch;
  constexpr int ES = CPLX ? 2 : 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t zb = i / MN, r = i - zb * MN;
    int64_t m = r / Nh, nh = r - m * Nh;
    double s = 0.0;
    for (int ks = 0; ks < p.splitk; ++ks)
      s += p.partial[((int64_t)ks * p.nbatch + zb) * MN + r];
    int64_t oa, om, on, ob, oc = 0;
    decode2(m, p.m, oa, om);
    decode2(CPLX ? (nh >> 1) : nh, p.n, ob, on);
    double *C = static_cast<double *>(p.C);
    if (p.dC) C = static_cast<double *>(p.dC[zb]);
    else {
      int64_t q = zb;
      for (int j = 0; j < p.b.n; ++j) {
        int64_t e = p.b.ext[j];
        int64_t qq = q / e;
        oc += (q - qq * e) * p.bsC[j];
        q = qq;
      }
    }
    double *dst = &C[(oc + om + on) * ES + (CPLX ? (nh & 1) : 0)];
    s *= p.alpha;
    if (p.beta != 0.0) s += p.beta * *dst;
    *dst = s;
  }
}

// Launch with programmatic stream serialisation: the CTAs of this kernel may
// become resident while the previous kernel in the stream drains (they build
// their offset tables, then block in griddepcontrol.wait until it has
// completed and flushed), which hides the launch latency between the
// back-to-back contractions of a tree / an MPS sweep.
template <typename Kern>
static int launch_pdl(Kern kern, dim3 grid, int threads, size_t smem, cudaStream_t st,
                      const ContractParams &p) {
  static const bool pdl = [] {
    const char *e = getenv("QB_PDL");
    return e ? atoi(e) != 0 : true;
  }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  if (pdl && cap == cudaStreamCaptureStatusNone) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  QB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
  QB_LAUNCH_CHECK();
  return 0;
}

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES,
          bool CPLX, int LA = -1, int LB = -1, int DBG = 0>
static int launch_cfg(const ContractParams &p, cudaStream_t st) {
  using Cfg = KernelCfg<BM, BN, BK, WARPS_M, WARPS_N, STAGES>;
  auto kern = contract_f64_kernel<BM, BN, BK, WARPS_M, WARPS_N, STAGES, CPLX, LA, LB, DBG>;
  static bool attr_set = false;
  if (!attr_set) {
    QB_CUDA_CHECK(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr_set = true;
  }
  if (p.nbatch > 65535 || p.splitk > 65535) {
    set_error("batch count %lld exceeds the launch limit", (long long)p.nbatch);
    return -100;
  }
  dim3 grid((unsigned)((int64_t)p.tiles_m * p.tiles_n), (unsigned)p.nbatch,
            (unsigned)p.splitk);
  return launch_pdl(kern, grid, Cfg::NT, Cfg::SMEM, st, p);
}

// layout of the operand tiles in shared memory ([row][k] = "k-major"); for
// real operands the big-tile kernels are specialised on it so that every
// fragment address is base + immediate
static inline int a_kmajor(const ContractParams &p) {
  return (p.vecA == 2) || (p.vecA == 0 && p.thrA == 1);
}
static inline int b_kmajor(const ContractParams &p) {
  return (p.vecB == 2) || (p.vecB == 0 && p.thrB == 1);
}
static inline bool layout_spec() {
  static const bool on = [] {
    const char *e = getenv("QB_LAYOUT_SPEC");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}

template <bool CPLX, int LA, int LB, int BK = 16, int STAGES = 4>
static int launch_streamk_l(const PairPlan &plan, cudaStream_t st) {
  using Cfg = KernelCfg<128, 128, BK, 4, 4, STAGES>;
  auto kern = contract_f64_streamk_kernel<128, 128, BK, 4, 4, STAGES, CPLX, LA, LB>;
  static bool attr_set = false;
  if (!attr_set) {
    QB_CUDA_CHECK(cudaFuncSetAttribute(
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr_set = true;
  }
  const int G = plan.streamk;
  if (!plan.flags_clean)
    QB_CUDA_CHECK(cudaMemsetAsync(plan.p.flags, 0, sizeof(int) * G, st));
  return launch_pdl(kern, dim3(G), Cfg::NT, Cfg::SMEM, st, plan.p);
}

template <bool CPLX>
static int launch_streamk(const PairPlan &plan, cudaStream_t st) {
  const bool bk32 = cfg0_bk() == 32;
  if (!CPLX && layout_spec()) {
    switch (a_kmajor(plan.p) * 2 + b_kmajor(plan.p) + (bk32 ? 4 : 0)) {
      case 0: return launch_streamk_l<false, 0, 0>(plan, st);
      case 1: return launch_streamk_l<false, 0, 1>(plan, st);
      case 2: return launch_streamk_l<false, 1, 0>(plan, st);
      case 3: return launch_streamk_l<false, 1, 1>(plan, st);
      case 4: return launch_streamk_l<false, 0, 0, 32, 3>(plan, st);
      case 5: return launch_streamk_l<false, 0, 1, 32, 3>(plan, st);
      case 6: return launch_streamk_l<false, 1, 0, 32, 3>(plan, st);
      default: return launch_streamk_l<false, 1, 1, 32, 3>(plan, st);
    }
  }
  if (bk32) return launch_streamk_l<CPLX, -1, -1, 32, 3>(plan, st);
  return launch_streamk_l<CPLX, -1, -1>(plan, st);
}

template <bool CPLX>
static int launch_cfg0(const ContractParams &p, cudaStream_t st) {
  if (!CPLX) {
    if (layout_spec()) {
      switch (a_kmajor(p) * 2 + b_kmajor(p) + (cfg0_bk() == 32 ? 4 : 0)) {
        case 0: return launch_cfg<128, 128, 16, 4, 4, 4, false, 0, 0>(p, st);
        case 1: return launch_cfg<128, 128, 16, 4, 4, 4, false, 0, 1>(p, st);
        case 2: return launch_cfg<128, 128, 16, 4, 4, 4, false, 1, 0>(p, st);
        case 3: return launch_cfg<128, 128, 16, 4, 4, 4, false, 1, 1>(p, st);
        case 4: return launch_cfg<128, 128, 32, 4, 4, 3, false, 0, 0>(p, st);
        case 5: return launch_cfg<128, 128, 32, 4, 4, 3, false, 0, 1>(p, st);
        case 6: return launch_cfg<128, 128, 32, 4, 4, 3, false, 1, 0>(p, st);
        default: return launch_cfg<128, 128, 32, 4, 4, 3, false, 1, 1>(p, st);
      }
    }
  }
  if (cfg0_bk() == 32) return launch_cfg<128, 128, 32, 4, 4, 3, CPLX>(p, st);
  return launch_cfg<128, 128, 16, 4, 4, 4, CPLX>(p, st);
}

template <bool CPLX>
static int launch_contract_t(const PairPlan &plan, cudaStream_t st) {
  const ContractParams &p = plan.p;
  if (plan.streamk > 0 && p.partial) return launch_streamk<CPLX>(plan, st);
  int rc;
  switch (plan.cfg) {
    case 0: rc = launch_cfg0<CPLX>(p, st); break;
    case 1: rc = launch_cfg<64, 64, 16, 2, 2, 4, CPLX>(p, st); break;
    case 2: rc = launch_cfg<128, 32, 16, 4, 1, 4, CPLX>(p, st); break;
    case 3: rc = launch_cfg<32, 128, 16, 1, 4, 4, CPLX>(p, st); break;
    default: rc = launch_cfg<32, 32, 16, 2, 1, 4, CPLX>(p, st); break;
  }
  if (rc) return rc;
  if (p.splitk > 1) {
    int64_t total = p.M * p.N * p.nbatch * (CPLX ? 2 : 1);
    int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
    splitk_reduce_f64_kernel<CPLX><<<blocks, 256, 0, st>>>(p);
    QB_LAUNCH_CHECK();
  }
  return 0;
}

int launch_contract_f64(const PairPlan &plan, cudaStream_t st) {
  return launch_contract_t<false>(plan, st);
}
int launch_contract_c128(const PairPlan &plan, cudaStream_t st) {
  return launch_contract_t<true>(plan, st);
}

}  // namespace qb
