// vtb200 — device input path: uint8 HWC source images -> normalised fp32 NCHW batch with the
// reference's mixup / cutmix / RandomErasing applied per sample, in ONE streaming kernel.
//
// Replaces, per sample, the CPU work of
//   mix_dataset.py:37-90   MixDataset.__getitem__  (Image.blend / paste on uint8, or mul/add_ / slice copy on tensors)
//   factory.py:163-174     ToTensor + Normalize(mean, std)
//   transforms.py:377-407  RandomErasing._erase (mode "pixel" / "const"; factory.py:178-182)
// The random DECISIONS (partner, ratio, boxes) stay on the host — they are drawn from Python's `random` in the
// reference's order by device_input.MixSampler and travel in a [B, 24] int32 table; this kernel is the byte work.
//
// HBM-bound: per output pixel 3 B (6 B with a partner) read, 12 B written; nothing is read twice.  The uint8 -> float
// normalisation is a 3 x 256 table built once per CTA in shared memory with IEEE intrinsics (bit-identical to
// torchvision's float32 div / sub / div whatever the compile flags).  One thread = 4 consecutive pixels of a row
// (3 x 32-bit loads per source, 3 x 128-bit streaming stores, one per colour plane) when W % 4 == 0; a scalar variant
// covers other widths.  Grid = a multiple of the SM count, grid-stride over (sample, pixel-tile) work items.
#include "../../include/vtb200.h"
#include "common.cuh"

namespace {

constexpr int IN_THREADS = 256;  // = the width of the normalisation table: thread t builds entry t
constexpr int IN_TABLE_COLS = 24;
constexpr uint32_t PHILOX_KEY1 = 0x7674B200u;

struct Row {
  int src1, src2, mode, domain;
  float w1, w2;
  int x1, y1, x2, y2;
  int a_top, a_left, a_h, a_w;
  int b_top, b_left, b_h, b_w;
  uint32_t seed_a, seed_b;
  int erase_mode;
};

__device__ __forceinline__ Row load_row(const int32_t* __restrict__ t) {
  Row r;
  r.src1 = __ldg(t + 0), r.src2 = __ldg(t + 1), r.mode = __ldg(t + 2), r.domain = __ldg(t + 3);
  r.w1 = __int_as_float(__ldg(t + 4)), r.w2 = __int_as_float(__ldg(t + 5));
  r.x1 = __ldg(t + 6), r.y1 = __ldg(t + 7), r.x2 = __ldg(t + 8), r.y2 = __ldg(t + 9);
  r.a_top = __ldg(t + 10), r.a_left = __ldg(t + 11), r.a_h = __ldg(t + 12), r.a_w = __ldg(t + 13);
  r.b_top = __ldg(t + 14), r.b_left = __ldg(t + 15), r.b_h = __ldg(t + 16), r.b_w = __ldg(t + 17);
  r.seed_a = (uint32_t)__ldg(t + 18), r.seed_b = (uint32_t)__ldg(t + 19);
  r.erase_mode = __ldg(t + 20);
  return r;
}

// Philox4x32-10 (Random123), counter (x, y, 0, 0), key (seed, PHILOX_KEY1)
__device__ __forceinline__ uint4 philox(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t c2 = 0, c3 = 0;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0, c1 = lo1, c2 = hi0 ^ c3 ^ k1, c3 = lo0;
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t x) {
  return __fmaf_rn((float)(x >> 8), 5.9604644775390625e-8f /* 2^-24 */, 2.98023223876953125e-8f /* 2^-25 */);
}
// three N(0, 1) values for the three channels of pixel (y, x)
__device__ __forceinline__ void noise3(uint32_t seed, int y, int x, float z[3]) {
  const uint4 r = philox((uint32_t)x, (uint32_t)y, seed, PHILOX_KEY1);
  const float two_pi = 6.2831853071795864769f;
  const float ra = sqrtf(-2.0f * logf(u01(r.x))), ta = two_pi * u01(r.y);
  const float rb = sqrtf(-2.0f * logf(u01(r.z))), tb = two_pi * u01(r.w);
  float s, c;
  sincosf(ta, &s, &c);
  z[0] = ra * c, z[1] = ra * s, z[2] = rb * cosf(tb);
}

__device__ __forceinline__ bool in_box(int y, int x, int top, int left, int h, int w) {
  return (unsigned)(y - top) < (unsigned)h && (unsigned)(x - left) < (unsigned)w;
}

// the value of one pixel (3 channels) of output sample `r` at (y, x); a[] / p[] = source bytes of img1 / partner
__device__ __forceinline__ void pixel(const Row& r, const float (*lut)[256], int y, int x, const uint32_t a[3],
                                      const uint32_t p[3], float v[3]) {
  const bool boxed = r.mode == 2 && y >= r.y1 && y < r.y2 && x >= r.x1 && x < r.x2;
  if (r.domain == 0) {
    // uint8 domain (PIL): blend / paste, then normalise, then erase the result (mix_before_aug = True order)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t m = a[c];
      if (r.mode == 1) {
        // ImagingBlend: (UINT8)((int)in1 + alpha * ((int)in2 - (int)in1)) in float, truncated
        const float d = (float)((int)p[c] - (int)a[c]);
        m = (uint32_t)(int)__fadd_rn((float)a[c], __fmul_rn(r.w1, d));
      } else if (boxed) {
        m = p[c];
      }
      v[c] = lut[c][m & 255u];
    }
    if (in_box(y, x, r.a_top, r.a_left, r.a_h, r.a_w)) {
      if (r.erase_mode == 1) noise3(r.seed_a, y, x, v);
      else v[0] = v[1] = v[2] = 0.f;
    }
    return;
  }
  // tensor domain: normalise + erase each source, then mul / add_ or slice copy (mix_before_aug = False order)
  float va[3], vb[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) va[c] = lut[c][a[c]], vb[c] = lut[c][p[c]];
  if (in_box(y, x, r.a_top, r.a_left, r.a_h, r.a_w)) {
    if (r.erase_mode == 1) noise3(r.seed_a, y, x, va);
    else va[0] = va[1] = va[2] = 0.f;
  }
  if (r.mode != 0 && in_box(y, x, r.b_top, r.b_left, r.b_h, r.b_w)) {
    if (r.erase_mode == 1) noise3(r.seed_b, y, x, vb);
    else vb[0] = vb[1] = vb[2] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (r.mode == 1) v[c] = __fmaf_rn(vb[c], r.w2, __fmul_rn(va[c], r.w1));  // img1.mul(ratio).add_(img2, alpha=1-ratio)
    else v[c] = boxed ? vb[c] : va[c];
  }
}

template <bool VEC>
__global__ void __launch_bounds__(IN_THREADS)
input_batch_kernel(const uint8_t* __restrict__ src, int n_src, const int32_t* __restrict__ table,
                   float* __restrict__ out, int B, int H, int W, float m0, float m1, float m2, float s0, float s1, float s2, int tiles_per_image) {
  __shared__ float lut[3][256];
  {
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
    const float v = __fdiv_rn((float)threadIdx.x, 255.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) lut[c][threadIdx.x] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
  }
  __syncthreads();
  const int64_t HW = (int64_t)H * W;
  constexpr int PPT = VEC ? 4 : 1;  // pixels per thread
  const int64_t n_items = (int64_t)B * tiles_per_image;
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = (int)(item / tiles_per_image), tile = (int)(item % tiles_per_image);
    const int64_t p0 = ((int64_t)tile * IN_THREADS + threadIdx.x) * PPT;
    if (p0 >= HW) continue;
    Row r = load_row(table + (int64_t)b * IN_TABLE_COLS);
    r.src1 = min(max(r.src1, 0), n_src - 1), r.src2 = min(max(r.src2, 0), n_src - 1);  // the host validates; stay in bounds
    const uint8_t* s1p = src + (int64_t)r.src1 * HW * 3 + p0 * 3;
    const uint8_t* s2p = src + (int64_t)r.src2 * HW * 3 + p0 * 3;
    const int y = (int)(p0 / W), x0 = (int)(p0 % W);
    float* o = out + (int64_t)b * 3 * HW + p0;
    if constexpr (VEC) {
      uint32_t wa[3], wb[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) wa[k] = __ldg(reinterpret_cast<const uint32_t*>(s1p) + k);
      if (r.mode != 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) wb[k] = __ldg(reinterpret_cast<const uint32_t*>(s2p) + k);
      } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) wb[k] = wa[k];
      }
      float res[3][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t a[3], p[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int k = j * 3 + c;  // byte index inside the 12-byte group (little endian words)
          a[c] = (wa[k >> 2] >> (8 * (k & 3))) & 255u;
          p[c] = (wb[k >> 2] >> (8 * (k & 3))) & 255u;
        }
        float v[3];
        pixel(r, lut, y, x0 + j, a, p, v);
#pragma unroll
        for (int c = 0; c < 3; ++c) res[c][j] = v[c];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        __stcs(reinterpret_cast<float4*>(o + c * HW), make_float4(res[c][0], res[c][1], res[c][2], res[c][3]));
    } else {
      uint32_t a[3], p[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) a[c] = __ldg(s1p + c);
#pragma unroll
      for (int c = 0; c < 3; ++c) p[c] = r.mode != 0 ? (uint32_t)__ldg(s2p + c) : a[c];
      float v[3];
      pixel(r, lut, y, x0, a, p, v);
#pragma unroll
      for (int c = 0; c < 3; ++c) __stcs(o + c * HW, v[c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ variant 2
// Same arithmetic and the same flat 1024-pixel tiles as variant 1, leaner per item.  Why: the first ncu capture of variant 1
// (profiles/r01_ncu_full_input_kernel.txt) shows 68 % of the issue slots busy, 464 warp instructions per 128 pixels, long_scoreboard as the top stall and 21 of 27
// global loads per thread spent on the decision row.  A first rewrite with row-shaped work items, the row staged in shared
// memory and two barriers per item was parity-green on the B200 but SLOWER (83 vs 72 us, profiles/
// r01_input_kernel_selftest_rowtiled_variant.log): with 4 CTAs per SM the barriers serialise the row -> source load chain.
// This variant keeps variant 1's barrier-free loop and only removes work:
//   * 32-bit index arithmetic (variant 1: three 64-bit divisions per thread and item);
//   * the decision row as six 16-byte read-only loads, specialised to the thread's image row (the y half of each box test
//     folded into a width) before the pixel loop;
//   * uint8 <-> float without the XU pipe: byte -> float by or-ing it into the mantissa of 2^23, the truncating
//     float -> byte of the PIL blend by a round-toward-zero add of 2^23;
//   * the erase noise (Philox + log + sincos) out of line, so the streaming path's register budget is not set by it.
__device__ __forceinline__ float byte_to_float(uint32_t b) { return __uint_as_float(0x4B000000u | b) - 8388608.0f; }

// out of line: erased pixels are the rare path, and Philox + log + sincos inlined four times per thread would set the
// register budget of the streaming path
__device__ __noinline__ float3 noise3_cold(uint32_t seed, int y, int x) {
  float z[3];
  noise3(seed, y, x, z);
  return make_float3(z[0], z[1], z[2]);  // by value: the callers' pixel arrays stay in registers
}

// The decision row specialised to one image row y: the y half of every box test is folded into the width (0 = miss).
struct RowY {
  int mode, domain, erase_mode;
  float w1, w2;
  int c_left, c_w, a_left, a_w, b_left, b_w;
  uint32_t seed_a, seed_b;
};
__device__ __forceinline__ bool in_span(int x, int left, int w) { return (unsigned)(x - left) < (unsigned)w; }

__device__ __forceinline__ void pixel2(const RowY& r, const float (*lut)[256], int y, int x, const uint32_t a[3],
                                       const uint32_t p[3], float v[3]) {
  const bool boxed = in_span(x, r.c_left, r.c_w);
  if (r.domain == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t m = a[c];
      if (r.mode == 1) {
        const float fa = byte_to_float(a[c]), d = __fsub_rn(byte_to_float(p[c]), fa);  // exact: small integers
        // 0 <= fa + alpha d <= 255 for alpha in [0, 1]: adding 2^23 toward zero leaves floor() in the low mantissa bits
        m = __float_as_uint(__fadd_rz(__fadd_rn(fa, __fmul_rn(r.w1, d)), 8388608.0f)) & 255u;
      } else if (boxed) {
        m = p[c];
      }
      v[c] = lut[c][m];
    }
    if (in_span(x, r.a_left, r.a_w)) {
      if (r.erase_mode == 1) { const float3 z = noise3_cold(r.seed_a, y, x); v[0] = z.x, v[1] = z.y, v[2] = z.z; }
      else v[0] = v[1] = v[2] = 0.f;
    }
    return;
  }
  float va[3], vb[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) va[c] = lut[c][a[c]], vb[c] = lut[c][p[c]];
  if (in_span(x, r.a_left, r.a_w)) {
    if (r.erase_mode == 1) { const float3 z = noise3_cold(r.seed_a, y, x); va[0] = z.x, va[1] = z.y, va[2] = z.z; }
    else va[0] = va[1] = va[2] = 0.f;
  }
  if (in_span(x, r.b_left, r.b_w)) {
    if (r.erase_mode == 1) { const float3 z = noise3_cold(r.seed_b, y, x); vb[0] = z.x, vb[1] = z.y, vb[2] = z.z; }
    else vb[0] = vb[1] = vb[2] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (r.mode == 1) v[c] = __fmaf_rn(vb[c], r.w2, __fmul_rn(va[c], r.w1));
    else v[c] = boxed ? vb[c] : va[c];
  }
}

// PPT pixels per thread: 1 (any width), 4 (W % 4 == 0: 3 x 32-bit loads per source, one 128-bit store per plane) or
// 8 (W % 8 == 0: 3 x 64-bit loads, two 128-bit stores per plane — variant 3: twice the bytes in flight per thread)
template <int PPT>
__global__ void __launch_bounds__(IN_THREADS)
input_batch_kernel2(const uint8_t* __restrict__ src, int n_src, const int32_t* __restrict__ table,
                    float* __restrict__ out, int B, int H, int W, float m0, float m1, float m2, float s0, float s1,
                    float s2, int tiles_per_image) {
  __shared__ float lut[3][256];
  {
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
    const float v = __fdiv_rn((float)threadIdx.x, 255.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) lut[c][threadIdx.x] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
  }
  __syncthreads();
  const uint32_t HW = (uint32_t)H * (uint32_t)W;  // < 2^31 (host check)
  const uint32_t n_items = (uint32_t)B * (uint32_t)tiles_per_image;
  for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const uint32_t b = item / (uint32_t)tiles_per_image, tile = item - b * (uint32_t)tiles_per_image;
    const uint32_t p0 = (tile * IN_THREADS + threadIdx.x) * PPT;
    if (p0 >= HW) continue;
    const int y = (int)(p0 / (uint32_t)W), x0 = (int)(p0 - (uint32_t)y * (uint32_t)W);
    // decision row: words 0..23 as six 16-byte loads (one address per warp)
    const int4* rp = reinterpret_cast<const int4*>(table + (int64_t)b * IN_TABLE_COLS);
    const int4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3), q4 = __ldg(rp + 4),
               q5 = __ldg(rp + 5);
    const int src1 = min(max(q0.x, 0), n_src - 1), src2 = min(max(q0.y, 0), n_src - 1);
    RowY r;
    r.mode = q0.z, r.domain = q0.w, r.erase_mode = q5.x;
    r.w1 = __int_as_float(q1.x), r.w2 = __int_as_float(q1.y);
    // cutmix box (x1 = q1.z, y1 = q1.w, x2 = q2.x, y2 = q2.y), half-open
    r.c_left = q1.z, r.c_w = (r.mode == 2 && y >= q1.w && y < q2.y) ? max(q2.x - q1.z, 0) : 0;
    // erase boxes (top, left, h, w): A = q2.z, q2.w, q3.x, q3.y;  B = q3.z, q3.w, q4.x, q4.y (only with a partner)
    r.a_left = q2.w, r.a_w = ((unsigned)(y - q2.z) < (unsigned)q3.x) ? q3.y : 0;
    r.b_left = q3.w, r.b_w = (r.mode != 0 && (unsigned)(y - q3.z) < (unsigned)q4.x) ? q4.y : 0;
    r.seed_a = (uint32_t)q4.z, r.seed_b = (uint32_t)q4.w;
    const uint8_t* s1p = src + ((int64_t)src1 * HW + p0) * 3;
    const uint8_t* s2p = src + ((int64_t)src2 * HW + p0) * 3;
    float* o = out + (int64_t)b * 3 * HW + p0;
    if constexpr (PPT > 1) {
      constexpr int NW = PPT * 3 / 4;  // 32-bit words of PPT packed RGB pixels
      uint32_t wa[NW], wb[NW];
      if constexpr (PPT == 8) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const uint2 t = __ldg(reinterpret_cast<const uint2*>(s1p) + k);
          wa[2 * k] = t.x, wa[2 * k + 1] = t.y;
        }
        if (r.mode != 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const uint2 t = __ldg(reinterpret_cast<const uint2*>(s2p) + k);
            wb[2 * k] = t.x, wb[2 * k + 1] = t.y;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < NW; ++k) wa[k] = __ldg(reinterpret_cast<const uint32_t*>(s1p) + k);
        if (r.mode != 0) {
#pragma unroll
          for (int k = 0; k < NW; ++k) wb[k] = __ldg(reinterpret_cast<const uint32_t*>(s2p) + k);
        }
      }
      if (r.mode == 0) {
#pragma unroll
        for (int k = 0; k < NW; ++k) wb[k] = wa[k];
      }
#pragma unroll
      for (int half = 0; half < PPT / 4; ++half) {  // four pixels at a time: one 128-bit store per colour plane
        float res[3][4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = half * 4 + jj;
          uint32_t a[3], p[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int k = j * 3 + c;  // byte index inside the packed group (little-endian words)
            a[c] = (wa[k >> 2] >> (8 * (k & 3))) & 255u;
            p[c] = (wb[k >> 2] >> (8 * (k & 3))) & 255u;
          }
          float v[3];
          pixel2(r, lut, y, x0 + j, a, p, v);
#pragma unroll
          for (int c = 0; c < 3; ++c) res[c][jj] = v[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
          __stcs(reinterpret_cast<float4*>(o + (int64_t)c * HW + half * 4),
                 make_float4(res[c][0], res[c][1], res[c][2], res[c][3]));
      }
    } else {
      uint32_t a[3], p[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) a[c] = __ldg(s1p + c);
#pragma unroll
      for (int c = 0; c < 3; ++c) p[c] = r.mode != 0 ? (uint32_t)__ldg(s2p + c) : a[c];
      float v[3];
      pixel2(r, lut, y, x0, a, p, v);
#pragma unroll
      for (int c = 0; c < 3; ++c) __stcs(o + (int64_t)c * HW, v[c]);
    }
  }
}

int g_input_variant = 2;  // 1: first kernel; 2: the lean one (default: 5-17 % faster on the B200, profiles/r02_input_variants.log); 3: lean, 8 pixels per thread (vtb_set_option("input_variant", v))

}  // namespace

void vtb_input_variant_set(int v) { g_input_variant = (v == 2 || v == 3) ? v : 1; }

extern "C" int vtb_input_batch(const uint8_t* src, int32_t n_src, const int32_t* table, int32_t batch, int32_t H, int32_t W,
                               const float* mean3, const float* std3, float* out, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VTB_CHECK(batch >= 0 && H > 0 && W > 0, -1, "vtb_input_batch: bad shape batch=%d H=%d W=%d", batch, H, W);
  if (batch == 0) return 0;  // an empty batch is legal (its table and output may be NULL)
  VTB_CHECK(src && table && out && mean3 && std3, -1, "vtb_input_batch: null pointer");
  VTB_CHECK(n_src > 0, -1, "vtb_input_batch: no source images (n_src=%d) for a batch of %d", n_src, batch);
  VTB_CHECK(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, -1, "vtb_input_batch: zero std");
  const int64_t HW = (int64_t)H * W;
  const bool vec = (W % 4 == 0) && (reinterpret_cast<uintptr_t>(src) % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const bool lean = g_input_variant >= 2 && reinterpret_cast<uintptr_t>(table) % 16 == 0 && HW < (1ll << 31);
  const bool wide = lean && g_input_variant == 3 && vec && W % 8 == 0 && reinterpret_cast<uintptr_t>(src) % 8 == 0;
  const int ppt = wide ? 8 : (vec ? 4 : 1);
  const int sms = vtb_num_sms() > 0 ? vtb_num_sms() : 148;  // vtb_init() not called yet: the B200 count
  const int64_t cap = (int64_t)sms * 8;                     // 8 resident CTAs of 256 threads per SM
  const int64_t tiles = (HW + (int64_t)IN_THREADS * ppt - 1) / ((int64_t)IN_THREADS * ppt);
  VTB_CHECK(tiles < (1ll << 30), -1, "vtb_input_batch: image too large");
  const int64_t items = (int64_t)batch * tiles;
  const int grid = (int)(items < cap ? items : cap);
  if (lean && items < (1ll << 31)) {
#define VTB_INPUT_LAUNCH2(P)                                                                                          \
  input_batch_kernel2<P><<<grid, IN_THREADS, 0, stream>>>(src, n_src, table, out, batch, H, W, mean3[0], mean3[1],    \
                                                          mean3[2], std3[0], std3[1], std3[2], (int)tiles)
    if (ppt == 8) VTB_INPUT_LAUNCH2(8);
    else if (ppt == 4) VTB_INPUT_LAUNCH2(4);
    else VTB_INPUT_LAUNCH2(1);
#undef VTB_INPUT_LAUNCH2
    VTB_LAUNCH_CHECK();
    return 0;
  }
  if (vec)
    input_batch_kernel<true><<<grid, IN_THREADS, 0, stream>>>(src, n_src, table, out, batch, H, W, mean3[0], mean3[1], mean3[2],
                                                              std3[0], std3[1], std3[2], (int)tiles);
  else
    input_batch_kernel<false><<<grid, IN_THREADS, 0, stream>>>(src, n_src, table, out, batch, H, W, mean3[0], mean3[1], mean3[2],
                                                               std3[0], std3[1], std3[2], (int)tiles);
  VTB_LAUNCH_CHECK();
  return 0;
}
