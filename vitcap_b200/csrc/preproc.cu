// Device-side head of the reference's TEST transform for a ragged batch of decoded 8-bit images:
//   Resize(resize_to, BICUBIC) on the shorter edge + CenterCrop(S)      (uni_pipeline.py:1233-1256; crop_pct 1.0 => resize_to == S)
// bit-identical to torchvision (output geometry) + Pillow ImagingResample (arithmetic): per-output-pixel windows of the Keys
// cubic (a = -0.5, support 2 * max(scale, 1)) computed in IEEE double precision with Pillow's operation order, normalised,
// converted to 22-bit fixed point; a horizontal pass into an 8-bit intermediate, then a vertical pass, int32 accumulation from
// 1 << 21 and a saturating shift. Only the cropped S x S window (and the source rows it needs) is ever computed. The channel
// order is untouched (the BGR->RGB flip commutes with per-channel resampling and is fused into vc_patchify_u8).
//
// HBM-bound byte work: per image ~ H*W*3 source bytes read once, rows*S*3 intermediate bytes written and read once, S*S*3 out.
#include "common.cuh"

namespace vc {

namespace {
constexpr int PREC = 22;                     // Pillow PRECISION_BITS = 32 - 8 - 2

struct Geometry { int in_size, out_size, off; };

// torchvision: shorter edge -> resize_to, longer edge -> int(resize_to * long / short); crop origin = round_half_even((full - S) / 2)
__device__ __host__ inline void resized_hw(int H, int W, int resize_to, int* nh, int* nw) {
  const int sh = W <= H ? W : H, lg = W <= H ? H : W;
  const int new_long = (int)((double)((long long)resize_to * lg) / (double)sh);
  *nw = (W <= H) ? resize_to : new_long;
  *nh = (W <= H) ? new_long : resize_to;
}

__device__ inline Geometry geometry(int H, int W, int resize_to, int S, int axis) {
  int nh, nw;
  resized_hw(H, W, resize_to, &nh, &nw);
  Geometry g;
  g.in_size = axis ? H : W;
  g.out_size = axis ? nh : nw;
  g.off = __double2int_rn((double)(g.out_size - S) / 2.0);      // exact halves round to even, as Python round()
  return g;
}

// Pillow bicubic_filter: explicit _rn intrinsics keep nvcc from contracting the products into FMAs
__device__ inline double cubic(double x) {
  x = fabs(x);
  if (x < 1.0) {
    double t = __dsub_rn(__dmul_rn(1.5, x), 2.5);               // (a + 2) x - (a + 3)
    t = __dmul_rn(__dmul_rn(t, x), x);
    return __dadd_rn(t, 1.0);
  }
  if (x < 2.0) {
    double t = __dmul_rn(__dsub_rn(x, 5.0), x);                  // ((x - 5) x + 8) x - 4
    t = __dmul_rn(__dadd_rn(t, 8.0), x);
    return __dmul_rn(__dsub_rn(t, 4.0), -0.5);
  }
  return 0.0;
}

// coefficient table of one (image, axis): int32 [kmax + 2][S]: row 0 = first source index, row 1 = tap count, rows 2.. = taps
__global__ void __launch_bounds__(128)
resize_coeff_kernel(const int* __restrict__ hw, int resize_to, int S, int kmax, int* __restrict__ coef) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int axis = blockIdx.y, b = blockIdx.z;
  if (i >= S) return;
  const Geometry g = geometry(hw[2 * b], hw[2 * b + 1], resize_to, S, axis);
  int* tab = coef + (size_t)(b * 2 + axis) * (kmax + 2) * S;
  const int xx = i + g.off;
  if (g.in_size == g.out_size) {              // Pillow skips the pass: identity window
    tab[i] = xx;
    tab[S + i] = 1;
    tab[2 * S + i] = 1 << PREC;
    for (int t = 1; t < kmax; ++t) tab[(2 + t) * S + i] = 0;
    return;
  }
  const double scale = (double)g.in_size / (double)g.out_size;
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, fscale);
  const double ss = 1.0 / fscale;
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  int lo = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (lo < 0) lo = 0;
  int hi = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (hi > g.in_size) hi = g.in_size;
  int n = hi - lo;
  if (n > kmax) n = kmax;                      // cannot happen with vc_resize_crop_plan's kmax; keeps the table in bounds
  double ww = 0.0;
  for (int x = 0; x < n; ++x)
    ww = __dadd_rn(ww, cubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + lo), center), 0.5), ss)));
  tab[i] = lo;
  tab[S + i] = n;
  for (int x = 0; x < kmax; ++x) {
    int q = 0;
    if (x < n) {
      double w = cubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + lo), center), 0.5), ss));
      if (ww != 0.0) w = w / ww;
      const double v = __dmul_rn(w, (double)(1 << PREC));
      q = (w < 0) ? (int)__dadd_rn(-0.5, v) : (int)__dadd_rn(0.5, v);
    }
    tab[(2 + x) * S + i] = q;
  }
}

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> PREC;                   // arithmetic shift (floor), then saturate
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: source rows [r0, r1) the vertical windows of the crop need, output columns = the S cropped columns
constexpr int RG = 8;                          // source rows per block (the taps are loaded once for all of them)
__global__ void __launch_bounds__(128)
resize_horizontal_kernel(const uint8_t* __restrict__ src, const long long* __restrict__ src_off, const int* __restrict__ hw,
                         int S, int kmax, const int* __restrict__ coef, uint8_t* __restrict__ tmp,
                         const long long* __restrict__ tmp_off) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int b = blockIdx.z;
  const int W = hw[2 * b + 1];
  const int* th = coef + (size_t)(b * 2) * (kmax + 2) * S;
  const int* tv = th + (size_t)(kmax + 2) * S;
  const int r0 = tv[0], r1 = tv[S - 1] + tv[S + S - 1];
  const int y0 = r0 + blockIdx.y * RG;
  if (i >= S || y0 >= r1) return;
  const int lo = th[i], n = th[S + i];
  int acc[RG][3];
#pragma unroll
  for (int r = 0; r < RG; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 1 << (PREC - 1);
  const uint8_t* p0 = src + src_off[b] + ((size_t)y0 * W + lo) * 3;
  const int rows = (r1 - y0) < RG ? (r1 - y0) : RG;
  for (int t = 0; t < n; ++t) {
    const int k = th[(2 + t) * S + i];
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      if (r < rows) {
        const uint8_t* p = p0 + ((size_t)r * W + t) * 3;
        acc[r][0] += (int)p[0] * k;
        acc[r][1] += (int)p[1] * k;
        acc[r][2] += (int)p[2] * k;
      }
    }
  }
  uint8_t* o = tmp + tmp_off[b] + ((size_t)y0 * S + i) * 3;
#pragma unroll
  for (int r = 0; r < RG; ++r) {
    if (r < rows) {
      o[(size_t)r * S * 3 + 0] = (uint8_t)clip8(acc[r][0]);
      o[(size_t)r * S * 3 + 1] = (uint8_t)clip8(acc[r][1]);
      o[(size_t)r * S * 3 + 2] = (uint8_t)clip8(acc[r][2]);
    }
  }
}

// vertical pass: the weights of an output row are the same for every byte of the row -> 4 bytes per thread, any channel mix
__global__ void __launch_bounds__(128)
resize_vertical_kernel(const uint8_t* __restrict__ tmp, const long long* __restrict__ tmp_off, int S, int kmax,
                       const int* __restrict__ coef, uint8_t* __restrict__ out) {
  const int q = blockIdx.x * 128 + threadIdx.x;          // 4-byte group inside the row
  const int j = blockIdx.y, b = blockIdx.z;
  const int row_words = S * 3 / 4;
  if (q >= row_words) return;
  const int* tv = coef + (size_t)(b * 2 + 1) * (kmax + 2) * S;
  const int lo = tv[j], n = tv[S + j];
  const uint8_t* base = tmp + tmp_off[b];
  int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0, a3 = a0;
  for (int t = 0; t < n; ++t) {
    const int k = tv[(2 + t) * S + j];
    const uint32_t w = *reinterpret_cast<const uint32_t*>(base + (size_t)(lo + t) * S * 3 + 4 * q);
    a0 += (int)(w & 255u) * k;
    a1 += (int)((w >> 8) & 255u) * k;
    a2 += (int)((w >> 16) & 255u) * k;
    a3 += (int)(w >> 24) * k;
  }
  const uint32_t r = (uint32_t)clip8(a0) | ((uint32_t)clip8(a1) << 8) | ((uint32_t)clip8(a2) << 16) | ((uint32_t)clip8(a3) << 24);
  *reinterpret_cast<uint32_t*>(out + ((size_t)b * S + j) * S * 3 + 4 * q) = r;
}
}  // namespace

// host-side planning (no device work): widest coefficient window of the batch, intermediate-buffer offsets, tallest image
int resize_crop_plan(const int* hw, int B, int resize_to, int S, int* kmax, long long* tmp_off, int* max_rows) {
  if (B <= 0 || S <= 0 || (S % 4) != 0 || resize_to < S) {
    set_last_error("resize_crop_plan: need B > 0, S %% 4 == 0 and resize_to >= S (crop_pct <= 1)");
    return VC_ERR_BAD_ARG;
  }
  int km = 1, mr = 0;
  long long off = 0;
  for (int b = 0; b < B; ++b) {
    const int H = hw[2 * b], W = hw[2 * b + 1];
    if (H <= 0 || W <= 0) { set_last_error("resize_crop_plan: image %d has size %d x %d", b, H, W); return VC_ERR_BAD_ARG; }
    int nh, nw;
    resized_hw(H, W, resize_to, &nh, &nw);
    if (nh < S || nw < S) { set_last_error("resize_crop_plan: image %d resizes below the crop", b); return VC_ERR_BAD_ARG; }
    const int in[2] = {W, H}, out[2] = {nw, nh};
    for (int a = 0; a < 2; ++a) {
      if (in[a] == out[a]) continue;
      const double scale = (double)in[a] / (double)out[a];
      const double support = 2.0 * (scale < 1.0 ? 1.0 : scale);
      const int ks = (int)ceil(support) * 2 + 1;
      if (ks > km) km = ks;
    }
    tmp_off[b] = off;
    off += (long long)H * S * 3;
    if (H > mr) mr = H;
  }
  tmp_off[B] = off;
  *kmax = km;
  *max_rows = mr;
  return VC_OK;
}

int resize_crop_u8(const uint8_t* src, const long long* src_off, const int* hw, int B, int resize_to, int S, int kmax, int max_rows,
                   int* coef, uint8_t* tmp, const long long* tmp_off, uint8_t* out, cudaStream_t s) {
  if (B <= 0 || B > 65535 || S <= 0 || (S % 4) != 0 || S > 65535 || resize_to < S || kmax < 1 || max_rows < 1 ||
      (reinterpret_cast<uintptr_t>(tmp) & 3) || (reinterpret_cast<uintptr_t>(out) & 3)) {
    set_last_error("resize_crop_u8: bad args");
    return VC_ERR_BAD_ARG;
  }
  const int cb = (S + 127) / 128;
  resize_coeff_kernel<<<dim3(cb, 2, B), 128, 0, s>>>(hw, resize_to, S, kmax, coef);
  int rc = check_launch("resize_coeff");
  if (rc) return rc;
  resize_horizontal_kernel<<<dim3(cb, (max_rows + RG - 1) / RG, B), 128, 0, s>>>(src, src_off, hw, S, kmax, coef, tmp, tmp_off);
  rc = check_launch("resize_horizontal");
  if (rc) return rc;
  resize_vertical_kernel<<<dim3((S * 3 / 4 + 127) / 128, S, B), 128, 0, s>>>(tmp, tmp_off, S, kmax, coef, out);
  return check_launch("resize_vertical");
}

}  // namespace vc
