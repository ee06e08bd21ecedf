// Precision conversion passes (HBM bound).  f32 / c64 tensors are contracted
// by widening them to f64 / c128 (exact), running the fp64 engines and
// rounding the result once -- every intermediate product/sum is carried in
// fp64, so results are at least as accurate as the reference's fp32 path.
// A native TF32x3 tcgen05 engine is the planned replacement.
#include "common.cuh"

namespace qb {

template <typename S, typename D>
__global__ void __launch_bounds__(256)
    convert_strided_kernel(const S *__restrict__ src, D *__restrict__ dst, int64_t n,
                           int rank, const int64_t *__restrict__ meta) {
  // meta: shape[rank], src_stride[rank] (elements); dst is contiguous
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i, off = 0;
    for (int d = rank - 1; d >= 0; --d) {
      const int64_t e = meta[d], q = r / e;
      off += (r - q * e) * meta[rank + d];
      r = q;
    }
    dst[i] = (D)src[off];
  }
}

template <typename S, typename D>
__global__ void __launch_bounds__(256)
    convert_contig_kernel(const S *__restrict__ src, D *__restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (D)src[i];
}

}  // namespace qb

using namespace qb;

extern "C" int qb_convert(int src_dtype, int dst_dtype, int64_t n, const void *src,
                          void *dst, void *stream) {
  // contiguous conversion between f32<->f64 and c64<->c128 (complex data is
  // converted as 2n reals)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n <= 0) return 0;
  int64_t nr = n;
  bool widen;
  if (src_dtype == QB_F32 && dst_dtype == QB_F64) widen = true;
  else if (src_dtype == QB_F64 && dst_dtype == QB_F32) widen = false;
  else if (src_dtype == QB_C64 && dst_dtype == QB_C128) { widen = true; nr = 2 * n; }
  else if (src_dtype == QB_C128 && dst_dtype == QB_C64) { widen = false; nr = 2 * n; }
  else {
    set_error("qb_convert: unsupported conversion %d -> %d", src_dtype, dst_dtype);
    return -1;
  }
  int64_t b = (nr + 255) / 256;
  if (b > (int64_t)sm_count() * 8) b = (int64_t)sm_count() * 8;
  if (widen)
    convert_contig_kernel<float, double><<<(unsigned)b, 256, 0, st>>>(
        (const float *)src, (double *)dst, nr);
  else
    convert_contig_kernel<double, float><<<(unsigned)b, 256, 0, st>>>(
        (const double *)src, (float *)dst, nr);
  QB_LAUNCH_CHECK();
  return 0;
}
