// Bandwidth-bound helper kernels of the caption path: patch extraction, token/context assembly,
// LayerNorm, row gathers, decode-step embedding. All use 128-bit global accesses and warp-shuffle
// reductions; one warp owns one row of the hidden dimension.
#include "pair.cuh"

namespace vc {

// ------------------------------------------------------------------------------------------
// patchify: image fp32 NCHW -> A[B*P, 3*p*p] (column order c,i,j == Conv2d weight.flatten(1));
// the conv of PatchEmbed (vision_transformer.py:267-275) then is a plain GEMM.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void patchify_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int HW, int p) {
  const int g = HW / p;                      // patches per side
  const int kdim = 3 * p * p;
  const int groups = kdim / 8;               // 8 consecutive j within one (c,i) row
  const size_t total = (size_t)B * g * g * groups;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int gi = (int)(idx % groups);
    const size_t prow = idx / groups;        // b*P + py*g + px
    const int px = (int)(prow % g);
    const int py = (int)((prow / g) % g);
    const int b = (int)(prow / ((size_t)g * g));
    const int k = gi * 8;
    const int c = k / (p * p), i = (k / p) % p, j = k % p;
    const float* src = img + (((size_t)b * 3 + c) * HW + (py * p + i)) * HW + px * p + j;
    float f[8];
    load8<float>(src, f);
    store8<T>(out + prow * kdim + k, f);
  }
}

int patchify(int out_bf16, const float* img, void* out, int B, int img_size, int patch, cudaStream_t s) {
  if (patch % 8 || img_size % patch) { set_last_error("patchify: patch %% 8 and img %% patch required"); return VC_ERR_BAD_ARG; }
  const size_t total = (size_t)B * (img_size / patch) * (img_size / patch) * (3 * patch * patch / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (out_bf16) patchify_kernel<bf16><<<blocks, 256, 0, s>>>(img, (bf16*)out, B, img_size, patch);
  else patchify_kernel<float><<<blocks, 256, 0, s>>>(img, (float*)out, B, img_size, patch);
  return check_launch("patchify");
}

// ------------------------------------------------------------------------------------------
// patchify_u8: the tail of the reference's test transform fused into the patch extraction. Input = what the host pipeline
// holds after BGR2RGB/resize/center-crop as 8-bit pixels, uint8 [B, S, S, 3] (HWC; channel order BGR or RGB); the kernel
// applies ToTensor (x / 255, HWC -> CHW) and Normalize(mean 0.5, std 0.5) = ((x / 255) - 0.5) / 0.5 with the same fp32
// operations in the same order as torchvision (uni_pipeline.py:1233-1256: BGR2RGB ... ToTensor, normalize) and writes the
// patch matrix of patchify(). The upload shrinks 4x (226 MB instead of 906 MB per 512 images).
// One thread = 8 consecutive pixels of one patch row, all three channels: 24 contiguous input bytes, three 8-element stores.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ img, T* __restrict__ out, int B, int HW, int p, int bgr) {
  const int g = HW / p;
  const int kdim = 3 * p * p;
  const int jg = p / 8;
  const size_t total = (size_t)B * g * g * p * jg;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % jg) * 8;
    const int i = (int)((idx / jg) % p);
    const size_t prow = idx / ((size_t)jg * p);
    const int px = (int)(prow % g);
    const int py = (int)((prow / g) % g);
    const int b = (int)(prow / ((size_t)g * g));
    const uint8_t* src = img + (((size_t)b * HW + (py * p + i)) * HW + (px * p + j)) * 3;
    uint2 w[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) w[q] = *reinterpret_cast<const uint2*>(src + 8 * q);
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(w);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cs = bgr ? 2 - c : c;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float v = __fdiv_rn((float)bytes[3 * e + cs], 255.0f);     // ToTensor
        f[e] = __fdiv_rn(v - 0.5f, 0.5f);                                   // Normalize(0.5, 0.5)
      }
      store8<T>(out + prow * kdim + (size_t)c * p * p + i * p + j, f);
    }
  }
}

int patchify_u8(int out_bf16, const uint8_t* img, void* out, int B, int img_size, int patch, int bgr, cudaStream_t s) {
  if (patch % 8 || img_size % patch || (reinterpret_cast<uintptr_t>(img) & 7)) {
    set_last_error("patchify_u8: patch %% 8, img %% patch and an 8-byte aligned image required");
    return VC_ERR_BAD_ARG;
  }
  const size_t total = (size_t)B * (img_size / patch) * (img_size / patch) * patch * (patch / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (out_bf16) patchify_u8_kernel<bf16><<<blocks, 256, 0, s>>>(img, (bf16*)out, B, img_size, patch, bgr);
  else patchify_u8_kernel<float><<<blocks, 256, 0, s>>>(img, (float*)out, B, img_size, patch, bgr);
  return check_launch("patchify_u8");
}

// ------------------------------------------------------------------------------------------
// assemble_tokens: x[b,0]=cls+pos[0]; x[b,1+p]=patch_out[b*P+p]+pos[1+p]  (vision_transformer.py:423-427)
// ------------------------------------------------------------------------------------------
__global__ void assemble_tokens_kernel(const float* __restrict__ patch_out, const float* __restrict__ cls,
                                       const float* __restrict__ pos, float* __restrict__ x, int B, int P, int H) {
  const int hv = H / 4;
  const size_t total = (size_t)B * (P + 1) * hv;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % hv) * 4;
    const size_t row = idx / hv;
    const int t = (int)(row % (P + 1));
    const int b = (int)(row / (P + 1));
    float4 a = (t == 0) ? *reinterpret_cast<const float4*>(cls + c)
                        : *reinterpret_cast<const float4*>(patch_out + ((size_t)b * P + t - 1) * H + c);
    float4 q = *reinterpret_cast<const float4*>(pos + (size_t)t * H + c);
    *reinterpret_cast<float4*>(x + row * H + c) = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
  }
}

int assemble_tokens(const float* patch_out, const float* cls, const float* pos, float* x, int B, int P, int H, cudaStream_t s) {
  const size_t total = (size_t)B * (P + 1) * (H / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  assemble_tokens_kernel<<<blocks, 256, 0, s>>>(patch_out, cls, pos, x, B, P, H);
  return check_launch("assemble_tokens");
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim (H % 128 == 0, H <= 1024); fp32 in; one warp per row.
// out_t: T copy (GEMM operand), out_f: optional fp32 copy (residual stream of the post-LN decoder).
// Two-pass mean/variance in registers (matches torch's fp32 LayerNorm to ~1 ulp).
// X3: out_t is the split-bf16 operand [hi | lo | hi] (3H columns, hi = bf16(y), lo = bf16(y - hi)) of the three-product
// tensor-core GEMM against [w_hi | w_hi | w_lo]; its first H columns are the plain bf16 copy.
// ------------------------------------------------------------------------------------------
template <typename T, bool X3 = false>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, T* __restrict__ out_t, int ld_t, float* __restrict__ out_f, int ld_f, int rows, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nv = H / 128;                    // float4 per lane
  float4 v[8];
  const float* ip = in + (size_t)warp * ld_in;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      v[i] = *reinterpret_cast<const float4*>(ip + (i * 32 + lane) * 4);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  const float mean = warp_sum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      const int c = (i * 32 + lane) * 4;
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 o = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                             (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      if (out_f) *reinterpret_cast<float4*>(out_f + (size_t)warp * ld_f + c) = o;
      if (out_t) {
        if (sizeof(T) == 4) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_t) + (size_t)warp * ld_t + c) = o;
        } else if (X3) {
          uint2 hi, lo;
          split_bf16x2(o.x, o.y, hi.x, lo.x);
          split_bf16x2(o.z, o.w, hi.y, lo.y);
          bf16* op = reinterpret_cast<bf16*>(out_t) + (size_t)warp * ld_t + c;
          *reinterpret_cast<uint2*>(op) = hi;
          *reinterpret_cast<uint2*>(op + H) = lo;
          *reinterpret_cast<uint2*>(op + 2 * H) = hi;
        } else {
          uint2 pk = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(out_t) + (size_t)warp * ld_t + c) = pk;
        }
      }
    }
}

int layernorm(int out_bf16, const float* in, int ld_in, const float* gamma, const float* beta, float eps, void* out_t, int ld_t,
              float* out_f, int ld_f, int rows, int H, cudaStream_t s) {
  if (H % 128 || H > 1024 || rows <= 0 || (ld_in % 4) || (ld_t % 4) || (out_f && (ld_f % 4))) {
    set_last_error("layernorm: need H %% 128 == 0, H <= 1024, pitches %% 4 (H=%d)", H);
    return VC_ERR_BAD_ARG;
  }
  const int blocks = (rows + 7) / 8;
  if (out_bf16 == 2) {
    if (!out_t || ld_t < 3 * H) { set_last_error("layernorm: the split-bf16 operand needs ld_t >= 3 H"); return VC_ERR_BAD_ARG; }
    launch_pdl(layernorm_kernel<bf16, true>, dim3(blocks), dim3(256), 0, s, in, ld_in, gamma, beta, eps, (bf16*)out_t, ld_t, out_f, ld_f, rows, H);
  } else if (out_bf16) launch_pdl(layernorm_kernel<bf16>, dim3(blocks), dim3(256), 0, s, in, ld_in, gamma, beta, eps, (bf16*)out_t, ld_t, out_f, ld_f, rows, H);
  else launch_pdl(layernorm_kernel<float>, dim3(blocks), dim3(256), 0, s, in, ld_in, gamma, beta, eps, (float*)out_t, ld_t, out_f, ld_f, rows, H);
  return check_launch("layernorm");
}

// ------------------------------------------------------------------------------------------
// split_bf16x3: fp32 rows [rows, K] -> split-bf16 operand [rows, 3K] = [hi | lo | hi] (see layernorm_kernel, X3).
// One thread per 8 consecutive elements: 32 B in, 3 x 16 B out.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float* __restrict__ in, int ld_in, bf16* __restrict__ out, int ld_out, int rows, int K) {
  pdl_launch_dependents();
  pdl_wait();
  const int per_row = K >> 3;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)rows * per_row) return;
  const int r = (int)(t / per_row), c = (int)(t % per_row) * 8;
  float f[8];
  load8<float>(in + (size_t)r * ld_in + c, f);
  uint4 hi, lo;
  split_bf16x2(f[0], f[1], hi.x, lo.x);
  split_bf16x2(f[2], f[3], hi.y, lo.y);
  split_bf16x2(f[4], f[5], hi.z, lo.z);
  split_bf16x2(f[6], f[7], hi.w, lo.w);
  bf16* op = out + (size_t)r * ld_out + c;
  *reinterpret_cast<uint4*>(op) = hi;
  *reinterpret_cast<uint4*>(op + K) = lo;
  *reinterpret_cast<uint4*>(op + 2 * K) = hi;
}

int split_bf16x3(const float* in, int ld_in, void* out, int ld_out, int rows, int K, cudaStream_t s) {
  if (rows <= 0 || K <= 0 || (K % 8) || (ld_in % 4) || (ld_out % 8) || ld_out < 3 * K ||
      (reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
    set_last_error("split_bf16x3: need K %% 8 == 0, ld_out >= 3 K, 16-byte aligned rows (K=%d)", K);
    return VC_ERR_BAD_ARG;
  }
  const long n = (long)rows * (K >> 3);
  launch_pdl(split_bf16x3_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, in, ld_in, (bf16*)out, ld_out, rows, K);
  return check_launch("split_bf16x3");
}

// ------------------------------------------------------------------------------------------
// gather_rows: out[r, :] = cast(in[r * row_stride + :H])  (e.g. the CLS row of every image)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void gather_rows_kernel(const float* __restrict__ in, size_t row_stride, T* __restrict__ out, int ld_out, int rows, int H) {
  const int hv = H / 8;
  const size_t total = (size_t)rows * hv;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % hv) * 8;
    const size_t r = idx / hv;
    float f[8];
    load8<float>(in + r * row_stride + c, f);
    store8<T>(out + r * ld_out + c, f);
  }
}

int gather_rows(int out_bf16, const float* in, size_t row_stride, void* out, int ld_out, int rows, int H, cudaStream_t s) {
  if (H % 8 || (row_stride % 4) || (ld_out % 8)) { set_last_error("gather_rows: alignment"); return VC_ERR_BAD_ARG; }
  const size_t total = (size_t)rows * (H / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (out_bf16) gather_rows_kernel<bf16><<<blocks, 256, 0, s>>>(in, row_stride, (bf16*)out, ld_out, rows, H);
  else gather_rows_kernel<float><<<blocks, 256, 0, s>>>(in, row_stride, (float*)out, ld_out, rows, H);
  return check_launch("gather_rows");
}

// ------------------------------------------------------------------------------------------
// assemble_ctx: ctx[b] = [ tag_stream[b,0] ; cap_stream[b,0..N-1] ]  (modeling_bert.py:1493) as fp32 residual
// copy + T operand copy. No LayerNorm / position / type embedding is applied to these rows.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void assemble_ctx_kernel(const float* __restrict__ cap, const float* __restrict__ tag, float* __restrict__ ctx_f,
                                    T* __restrict__ ctx_t, int B, int N, int H, int Cp) {
  const int hv = H / 8;
  const size_t total = (size_t)B * (N + 1) * hv;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % hv) * 8;
    const size_t row = idx / hv;
    const int t = (int)(row % (N + 1));
    const size_t b = row / (N + 1);
    const float* src = (t == 0) ? tag + (b * N) * H + c : cap + (b * N + t - 1) * H + c;
    float f[8];
    load8<float>(src, f);
    const size_t orow = b * Cp + t;                // Cp >= N + 1 context rows are allocated per image
    store8<float>(ctx_f + orow * H + c, f);
    if (sizeof(T) == 2) store8<T>(ctx_t + orow * H + c, f);
  }
}

int assemble_ctx(int out_bf16, const float* cap, const float* tag, float* ctx_f, void* ctx_t, int B, int N, int H, int Cp,
                 cudaStream_t s) {
  if (H % 8 || Cp < N + 1) { set_last_error("assemble_ctx: H %% 8, rows per image >= N + 1"); return VC_ERR_BAD_ARG; }
  const size_t total = (size_t)B * (N + 1) * (H / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (out_bf16) assemble_ctx_kernel<bf16><<<blocks, 256, 0, s>>>(cap, tag, ctx_f, (bf16*)ctx_t, B, N, H, Cp);
  else assemble_ctx_kernel<float><<<blocks, 256, 0, s>>>(cap, tag, ctx_f, (float*)ctx_t, B, N, H, Cp);
  return check_launch("assemble_ctx");
}

// ------------------------------------------------------------------------------------------
// embed_ln: decode-step text embedding (BertEmbeddings.forward, modeling_bert.py:222-237) for the two rows of
// every sequence: [last generated token @ pos cur_len-1, MASK @ pos cur_len]; token type 0.
// ids: int32 [R, max_len] running token buffer. One warp per output row.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
embed_ln_kernel(const int* __restrict__ ids, int max_len, int cur_len, int mask_id, const float* __restrict__ word,
                const float* __restrict__ pos, const float* __restrict__ type0, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, float* __restrict__ out_f, T* __restrict__ out_t, int R, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= 2 * R) return;
  const int r = warp >> 1, which = warp & 1;
  const int tok = which ? mask_id : ids[(size_t)r * max_len + cur_len - 1];
  const int p = which ? cur_len : cur_len - 1;
  const int nv = H / 128;
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      const int c = (i * 32 + lane) * 4;
      float4 a = __ldg(reinterpret_cast<const float4*>(word + (size_t)tok * H + c));
      float4 b = __ldg(reinterpret_cast<const float4*>(pos + (size_t)p * H + c));
      float4 t = __ldg(reinterpret_cast<const float4*>(type0 + c));
      v[i] = make_float4(a.x + b.x + t.x, a.y + b.y + t.y, a.z + b.z + t.z, a.w + b.w + t.w);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  const float mean = warp_sum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) {
      const int c = (i * 32 + lane) * 4;
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 o = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                             (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      *reinterpret_cast<float4*>(out_f + (size_t)warp * H + c) = o;
      if (sizeof(T) == 2) {
        uint2 pk = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(out_t) + (size_t)warp * H + c) = pk;
      }
    }
}

int embed_ln(int out_bf16, const int* ids, int max_len, int cur_len, int mask_id, const float* word, const float* pos,
             const float* type0, const float* gamma, const float* beta, float eps, float* out_f, void* out_t, int R, int H,
             cudaStream_t s) {
  if (H % 128 || H > 1024 || cur_len < 1 || cur_len >= max_len) { set_last_error("embed_ln: bad args"); return VC_ERR_BAD_ARG; }
  const int blocks = (2 * R + 7) / 8;
  if (out_bf16) launch_pdl(embed_ln_kernel<bf16>, dim3(blocks), dim3(256), 0, s, ids, max_len, cur_len, mask_id, word, pos, type0, gamma, beta, eps, out_f, (bf16*)out_t, R, H);
  else launch_pdl(embed_ln_kernel<float>, dim3(blocks), dim3(256), 0, s, ids, max_len, cur_len, mask_id, word, pos, type0, gamma, beta, eps, out_f, (float*)out_t, R, H);
  return check_launch("embed_ln");
}

// ------------------------------------------------------------------------------------------
// label_rows: the od/tag label rows of the context when the caller's mask makes them visible.
//   row (b, i) = E_word[tag_i]                                   recipe 0 ('raw', modeling_bert.py:1447-1470)
//              = LN(E_word[tag_i] + E_pos[pos0 + i] + E_type[0])  recipe 1 ('ln', encode_tag_to_embedding :1381-1406)
// with tag_i = i-th predicted concept of image b and the last slot forced to [SEP] (:1447 / :1477). Written as the fp32
// residual copy and the T operand copy to context rows row0 .. row0 + K - 1 of image b (Cp rows per image). Warp per row.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
label_rows_kernel(const int* __restrict__ tag_idx, int K, int sep_id, int recipe_ln, int pos0, const float* __restrict__ word,
                  const float* __restrict__ pos, const float* __restrict__ type0, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, float* __restrict__ ctx_f, T* __restrict__ ctx_t, int B, int Cp,
                  int row0, int H) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * K) return;
  const int b = warp / K, i = warp - b * K;
  const int tok = (i == K - 1) ? sep_id : tag_idx[(size_t)b * K + i];
  const size_t orow = (size_t)b * Cp + row0 + i;
  const int nv = H / 128;
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const int c = (j * 32 + lane) * 4;
      v[j] = __ldg(reinterpret_cast<const float4*>(word + (size_t)tok * H + c));
      if (recipe_ln) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (size_t)(pos0 + i) * H + c));
        const float4 t = __ldg(reinterpret_cast<const float4*>(type0 + c));
        v[j] = make_float4(v[j].x + p.x + t.x, v[j].y + p.y + t.y, v[j].z + p.z + t.z, v[j].w + p.w + t.w);
      }
      s += v[j].x + v[j].y + v[j].z + v[j].w;
    }
  float mean = 0.f, rstd = 1.f;
  if (recipe_ln) {
    mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nv) {
        const float a = v[j].x - mean, bb = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        q += a * a + bb * bb + c * c + d * d;
      }
    rstd = rsqrtf(warp_sum(q) / (float)H + eps);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const int c = (j * 32 + lane) * 4;
      float4 o = v[j];
      if (recipe_ln) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
        o = make_float4((v[j].x - mean) * rstd * g.x + be.x, (v[j].y - mean) * rstd * g.y + be.y,
                        (v[j].z - mean) * rstd * g.z + be.z, (v[j].w - mean) * rstd * g.w + be.w);
      }
      *reinterpret_cast<float4*>(ctx_f + orow * H + c) = o;
      if (sizeof(T) == 2) {
        const uint2 pk = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ctx_t) + orow * H + c) = pk;
      }
    }
}

int label_rows(int out_bf16, const int* tag_idx, int K, int sep_id, int recipe_ln, int pos0, const float* word, const float* pos,
               const float* type0, const float* gamma, const float* beta, float eps, float* ctx_f, void* ctx_t, int B, int Cp,
               int row0, int H, cudaStream_t s) {
  if (H % 128 || H > 1024 || K < 1 || B < 1 || row0 + K > Cp) { set_last_error("label_rows: bad args"); return VC_ERR_BAD_ARG; }
  const int blocks = (B * K + 7) / 8;
  if (out_bf16) label_rows_kernel<bf16><<<blocks, 256, 0, s>>>(tag_idx, K, sep_id, recipe_ln, pos0, word, pos, type0, gamma, beta, eps, ctx_f, (bf16*)ctx_t, B, Cp, row0, H);
  else label_rows_kernel<float><<<blocks, 256, 0, s>>>(tag_idx, K, sep_id, recipe_ln, pos0, word, pos, type0, gamma, beta, eps, ctx_f, (float*)ctx_t, B, Cp, row0, H);
  return check_launch("label_rows");
}

}  // namespace vc
