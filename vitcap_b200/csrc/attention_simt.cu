// CUDA-core flash-style self-attention over a packed QKV buffer (exact mode, and cross-check of the
// tcgen05 attention kernel). softmax(Q K^T * scale) V per (image, head); scores never leave the SM.
//   qkv : [B, N, 3*H] (q | k | v, each H = heads*64 wide)   (vision_transformer.py:174-200,
//   out : [B, N, H]                                           modeling_bert.py:303-340 with an all-visible mask)
#include "common.cuh"

namespace vc {

template <typename T>
__global__ void __launch_bounds__(256)
attention_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N, int H, float scale) {
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
      if (k0 + lane >= N) s0 = -INFINITY;
      if (k0 + lane + 32 >= N) s1 = -INFINITY;
      const float mn = fmaxf(m[qi], warp_max(fmaxf(s0, s1)));
      const float corr = expf(m[qi] - mn);
      const float p0 = expf(s0 - mn), p1 = expf(s1 - mn);
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

int attention_simt(int is_bf16, const void* qkv, void* out, int B, int N, int heads, float scale, cudaStream_t s) {
  if (B <= 0 || N <= 0 || heads <= 0) { set_last_error("attention_simt: bad args"); return VC_ERR_BAD_ARG; }
  dim3 grid((N + 31) / 32, heads, B);
  const int H = heads * 64;
  if (is_bf16) attention_simt_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)qkv, (bf16*)out, N, H, scale);
  else attention_simt_kernel<float><<<grid, 256, 0, s>>>((const float*)qkv, (float*)out, N, H, scale);
  return check_launch("attention_simt");
}

}  // namespace vc
