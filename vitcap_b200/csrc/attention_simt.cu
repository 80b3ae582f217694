// CUDA-core flash-style self-attention over a packed QKV buffer (exact mode, and cross-check of the
// tcgen05 attention kernel). softmax(Q K^T * scale) V per (image, head); scores never leave the SM.
//   qkv : [B, N, 3*H] (q | k | v, each H = heads*64 wide)   (vision_transformer.py:174-200,
//   out : [B, N, H]                                           modeling_bert.py:303-340 with an all-visible mask)
#include "common.cuh"

namespace vc {

template <typename T>
__global__ void __launch_bounds__(256)
attention_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N, int H, float scale, int n_base,
                      const int* __restrict__ n_extra) {
  constexpr int D = 64, QB = 32, KB = 64;
  __shared__ float Qs[QB][D];
  __shared__ float Ks[KB][D + 1];
  __shared__ float Vs[KB][D];
  __shared__ float Ps[8][KB];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t ld = 3 * (size_t)H;
  const T* base = qkv + (size_t)b * N * ld;

  // stage Q (32 x 64): thread -> (row = tid/8, 8 consecutive dims)
  {
    const int r = tid >> 3, c = (tid & 7) * 8;
    float f[8];
    if (q0 + r < N) load8<T>(base + (size_t)(q0 + r) * ld + h * D + c, f);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) Qs[r][c + i] = f[i] * scale;
  }
  float m[4], l[4], o0[4], o1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; o0[i] = 0.f; o1[i] = 0.f; }
  // label-region mask (n_extra != NULL): rows below n_base see the keys below n_base, later rows n_extra[b] more
  const int lab_lim = n_extra ? n_base + n_extra[b] : N;

  for (int k0 = 0; k0 < N; k0 += KB) {
    __syncthreads();
    for (int t = tid; t < KB * 8; t += 256) {
      const int r = t >> 3, c = (t & 7) * 8;
      float fk[8], fv[8];
      if (k0 + r < N) {
        load8<T>(base + (size_t)(k0 + r) * ld + H + h * D + c, fk);
        load8<T>(base + (size_t)(k0 + r) * ld + 2 * H + h * D + c, fv);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { fk[i] = 0.f; fv[i] = 0.f; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { Ks[r][c + i] = fk[i]; Vs[r][c + i] = fv[i]; }
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const int q = warp * 4 + qi;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 16
      for (int d = 0; d < D; ++d) {
        const float qv = Qs[q][d];
        s0 = fmaf(qv, Ks[lane][d], s0);
        s1 = fmaf(qv, Ks[lane + 32][d], s1);
      }
      const int lim = (n_extra && q0 + q < n_base) ? n_base : lab_lim;
      if (k0 + lane >= lim) s0 = -INFINITY;
      if (k0 + lane + 32 >= lim) s1 = -INFINITY;
      const float mn = fmaxf(m[qi], warp_max(fmaxf(s0, s1)));
      const float corr = (m[qi] == -INFINITY) ? 0.f : expf(m[qi] - mn);
      const float p0 = (s0 == -INFINITY) ? 0.f : expf(s0 - mn), p1 = (s1 == -INFINITY) ? 0.f : expf(s1 - mn);
      l[qi] = l[qi] * corr + warp_sum(p0 + p1);
      m[qi] = mn;
      __syncwarp();
      Ps[warp][lane] = p0;
      Ps[warp][lane + 32] = p1;
      __syncwarp();
      float a0 = o0[qi] * corr, a1 = o1[qi] * corr;
#pragma unroll 16
      for (int k = 0; k < KB; ++k) {
        const float p = Ps[warp][k];
        a0 = fmaf(p, Vs[k][lane], a0);
        a1 = fmaf(p, Vs[k][lane + 32], a1);
      }
      o0[qi] = a0; o1[qi] = a1;
    }
  }
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int q = q0 + warp * 4 + qi;
    if (q < N) {
      const float inv = 1.f / l[qi];
      T* op = out + ((size_t)b * N + q) * H + h * D;
      op[lane] = from_f32<T>(o0[qi] * inv);
      op[lane + 32] = from_f32<T>(o1[qi] * inv);
    }
  }
}

int attention_simt(int is_bf16, const void* qkv, void* out, int B, int N, int heads, float scale, int n_base, const int* n_extra,
                   cudaStream_t s) {
  if (B <= 0 || N <= 0 || heads <= 0 || (n_extra && (n_base < 1 || n_base > N))) {
    set_last_error("attention_simt: bad args"); return VC_ERR_BAD_ARG;
  }
  dim3 grid((N + 31) / 32, heads, B);
  const int H = heads * 64;
  if (is_bf16) attention_simt_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)qkv, (bf16*)out, N, H, scale, n_base, n_extra);
  else attention_simt_kernel<float><<<grid, 256, 0, s>>>((const float*)qkv, (float*)out, N, H, scale, n_base, n_extra);
  return check_launch("attention_simt");
}

}  // namespace vc

// ------------------------------------------------------------------------------------------
// One query row per image against all N keys of the packed qkv buffer: the LAST block of the concept (tag) branch.
// Only the CLS row of that block's output is ever consumed -- by the pooler / tag head (modeling_bert.py:1424-1425) and as
// the tag token prepended to the visual context (modeling_bert.py:1493) -- so the block computes K, V for all rows but
// Q, attention, proj and the MLP for row 0 only (identical mathematics for that row, 85 % of the block's FLOPs skipped).
// HBM-bound: reads K and V of one (image, head) once; 8 lanes share a 64-dim row (16-byte loads), 4 warps split the keys.
// ------------------------------------------------------------------------------------------
namespace vc {

// 8 consecutive elements as loaded (16 B of bf16 / 32 B of fp32): kept raw until use so that many loads fit in registers
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void get(float* f) const {
    unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]); unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void get(float* f) const { f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w; }
};

template <typename T>
__global__ void __launch_bounds__(128)
cls_attention_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ qkv, T* __restrict__ out, int ldo, int N, int H,
                     float scale) {
  constexpr int D = 64, U = 8;          // keys per lane group and iteration: 16 vector loads in flight per thread
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, part = lane & 7;
  const size_t ld = 3 * (size_t)H;
  float qf[8], o[8];
  load8<T>(q + (size_t)b * ldq + h * D + part * 8, qf);
#pragma unroll
  for (int d = 0; d < 8; ++d) { qf[d] *= scale; o[d] = 0.f; }
  float m = -1e30f, l = 0.f;
  const T* kbase = qkv + (size_t)b * N * ld + H + h * D + part * 8;
  const T* vbase = kbase + H;
  for (int kb = warp * 4; kb < N; kb += 16 * U) {        // warp-uniform trip count (shuffles inside)
    const int k0 = kb + grp;
    Raw8<T> kr[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + 16 * u;
      if (k < N) {
        kr[u].load(kbase + (size_t)k * ld);
        vr[u].load(vbase + (size_t)k * ld);
      } else {
        kr[u].zero();
        vr[u].zero();
      }
    }
    float s[U];
    float mx = m;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float kf[8];
      kr[u].get(kf);
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) a = fmaf(qf[d], kf[d], a);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      s[u] = (k0 + 16 * u < N) ? a : -1e30f;
      mx = fmaxf(mx, s[u]);
    }
    const float corr = expf(m - mx);
    m = mx;
    l *= corr;
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] *= corr;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float p = (k0 + 16 * u < N) ? expf(s[u] - mx) : 0.f;
      l += p;
      float vf[8];
      vr[u].get(vf);
#pragma unroll
      for (int d = 0; d < 8; ++d) o[d] = fmaf(p, vf[d], o[d]);
    }
  }
  // merge the 4 lane groups of the warp, then the 4 warps
#pragma unroll
  for (int x = 8; x <= 16; x <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, x);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, x);
    const float mx = fmaxf(m, m2);
    const float c1 = expf(m - mx), c2 = expf(m2 - mx);
    l = l * c1 + l2 * c2;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const float o2 = __shfl_xor_sync(0xffffffffu, o[d], x);
      o[d] = o[d] * c1 + o2 * c2;
    }
    m = mx;
  }
  __shared__ float sm_m[4], sm_l[4], sm_o[4][D];
  if (grp == 0) {
    if (part == 0) { sm_m[warp] = m; sm_l[warp] = l; }
#pragma unroll
    for (int d = 0; d < 8; ++d) sm_o[warp][part * 8 + d] = o[d];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int pt = threadIdx.x;
    float mx = sm_m[0];
#pragma unroll
    for (int w = 1; w < 4; ++w) mx = fmaxf(mx, sm_m[w]);
    float lt = 0.f, acc[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float c = expf(sm_m[w] - mx);
      lt += sm_l[w] * c;
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] += sm_o[w][pt * 8 + d] * c;
    }
    const float inv = 1.f / lt;
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] *= inv;
    store8<T>(out + (size_t)b * ldo + h * D + pt * 8, acc);
  }
}

// q [B, heads*64] (row pitch ldq), qkv [B, N, 3*heads*64], out [B, heads*64] (row pitch ldo)
int cls_attention(int is_bf16, const void* q, int ldq, const void* qkv, void* out, int ldo, int B, int N, int heads, float scale,
                  cudaStream_t s) {
  if (B <= 0 || N <= 0 || heads <= 0 || B > 65535 || (ldq % 8) || (ldo % 8)) { set_last_error("cls_attention: bad args"); return VC_ERR_BAD_ARG; }
  const int H = heads * 64;
  if (is_bf16)
    cls_attention_kernel<bf16><<<dim3(heads, B), 128, 0, s>>>((const bf16*)q, ldq, (const bf16*)qkv, (bf16*)out, ldo, N, H, scale);
  else
    cls_attention_kernel<float><<<dim3(heads, B), 128, 0, s>>>((const float*)q, ldq, (const float*)qkv, (float*)out, ldo, N, H, scale);
  return check_launch("cls_attention");
}

}  // namespace vc
