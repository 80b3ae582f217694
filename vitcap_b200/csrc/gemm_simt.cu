// CUDA-core GEMM with fp32 accumulation: out[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ resid).
// Used for (a) the exact mode (fp32 operands; token IDs / tag indices bit-identical to the fp32
// reference need fp32 operands, SURVEY.md section 7 "exactness modes") and (b) as an independent
// on-device cross-check of the tcgen05 kernel in tests. 128x128x16 tiles, 8x8 outputs per thread.
#include "common.cuh"

namespace vc {

enum { SACT_NONE = 0, SACT_GELU = 1, SACT_TANH = 2 };

template <typename TI, typename TO, int ACT, bool RESID>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TI* __restrict__ A, int lda, const TI* __restrict__ W, int ldw, const float* __restrict__ bias,
                 TO* __restrict__ out, int ldo, const float* resid, int ldr, int M, int N, int K) {
  constexpr int BM = 128, BN = 128, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 1, lk = (tid & 1) * 8;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    float fa[8], fb[8];
    if (m0 + lrow < M) load8<TI>(A + (size_t)(m0 + lrow) * lda + k0 + lk, fa);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) fa[i] = 0.f;
    }
    if (n0 + lrow < N) load8<TI>(W + (size_t)(n0 + lrow) * ldw + k0 + lk, fb);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) fb[i] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) { As[lk + i][lrow] = fa[i]; Bs[lk + i][lrow] = fb[i]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  const int col0 = n0 + tx * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty * 8 + i;
    if (row >= M) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = col0 + j;
      float x = acc[i][j];
      if (col < N) {
        if (bias != nullptr) x += __ldg(bias + col);
        if (ACT == SACT_GELU) x = gelu_erf(x);
        else if (ACT == SACT_TANH) x = tanhf(x);
        if (RESID) x += resid[(size_t)row * ldr + col];
      }
      v[j] = x;
    }
    TO* op = out + (size_t)row * ldo + col0;
    if (col0 + 8 <= N && ((reinterpret_cast<uintptr_t>(op) & (sizeof(TO) * 8 - 1)) == 0)) {
      store8<TO>(op, v);
    } else {
      for (int j = 0; j < 8; ++j)
        if (col0 + j < N) op[j] = from_f32<TO>(v[j]);
    }
  }
}

template <typename TI, typename TO>
static int launch_simt(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int act,
                       const float* resid, int ldr, int M, int N, int K, cudaStream_t s) {
  dim3 grid((N + 127) / 128, (M + 127) / 128);
  const TI* a = reinterpret_cast<const TI*>(A);
  const TI* w = reinterpret_cast<const TI*>(W);
  TO* o = reinterpret_cast<TO*>(out);
#define VC_LAUNCH(ACTV, RES) gemm_simt_kernel<TI, TO, ACTV, RES><<<grid, 256, 0, s>>>(a, lda, w, ldw, bias, o, ldo, resid, ldr, M, N, K)
  if (resid) {
    if (act == SACT_NONE) VC_LAUNCH(SACT_NONE, true);
    else if (act == SACT_GELU) VC_LAUNCH(SACT_GELU, true);
    else VC_LAUNCH(SACT_TANH, true);
  } else {
    if (act == SACT_NONE) VC_LAUNCH(SACT_NONE, false);
    else if (act == SACT_GELU) VC_LAUNCH(SACT_GELU, false);
    else VC_LAUNCH(SACT_TANH, false);
  }
#undef VC_LAUNCH
  return check_launch("gemm_simt");
}

// in_bf16: operands bf16 (else fp32); out_f32: output fp32 (else bf16)
int gemm_simt(int in_bf16, const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo,
              int out_f32, int act, const float* resid, int ldr, int M, int N, int K, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 16) || (lda % 8) || (ldw % 8) || act < 0 || act > 2) {
    set_last_error("gemm_simt: bad shape M=%d N=%d K=%d lda=%d ldw=%d act=%d (K%%16, ld%%8 required)", M, N, K, lda, ldw, act);
    return VC_ERR_BAD_ARG;
  }
  if (in_bf16) {
    if (out_f32) return launch_simt<bf16, float>(A, lda, W, ldw, bias, out, ldo, act, resid, ldr, M, N, K, s);
    return launch_simt<bf16, bf16>(A, lda, W, ldw, bias, out, ldo, act, resid, ldr, M, N, K, s);
  }
  if (out_f32) return launch_simt<float, float>(A, lda, W, ldw, bias, out, ldo, act, resid, ldr, M, N, K, s);
  return launch_simt<float, bf16>(A, lda, W, ldw, bias, out, ldo, act, resid, ldr, M, N, K, s);
}

}  // namespace vc
