// fcb_umma_v1.cu -- FIRST-GENERATION "umma_i8" kernel (one TMA box per filter tap, M = 128 pixels).  NOT part of the product
// library: it is compiled only into experiment builds (-DFCB_EXPERIMENT, tools/libfinnconv_exp.so) where FCB_UMMA_V1=1 selects it
// as an independent cross-check of the resident-planes kernel (fcb_umma2.cu).  Host-side plan glue lives in fcb_plan.cu.
//
//   D[pixel][ch] = sum_k A[pixel][k] * W[ch][k],  k = (ky*Kx + kx)*C + c   (mvau.hpp:122-178,
//   window order slidingwindow.h:1302-1313), followed by the fused activation stage.
//
// What replaces what:
//   FMPadding_nonsquare (streamtools.h:361-406)        -> TMA out-of-bounds zero fill (negative / past-the-end box coordinates)
//   ConvolutionInputGenerator_NonSquare + the stride    -> one TMA box per filter tap: a BH x BW patch of output pixels is a BH x BW
//   decimation loop (conv_nonsquare_top.cpp:238-259)       box of input pixels; for stride 2 the NHWC image is viewed as the 5-D
//                                                          tensor [n][y/2][y%2][x/2][(x%2)*C + c] so a fixed tap is again a dense box
//   zero insertion of deconv522 (:109-156)              -> 4 output phases, each a stride-1 conv over the ORIGINAL input with the
//                                                          3x3 / 2x3 / 3x2 / 2x2 taps that hit non-zero samples
//   Matrix_Vector_Activate_Batch                         -> tcgen05.mma.cta_group::1.kind::i8, M = 128 pixels, N = OFM, K = 32 per instr
//   PassThrough/bias+ReLU/Thresholds (+ max pool)        -> epilogue on the TMEM accumulators (fcb_epilogue.cuh)
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue (TMEM lane quarter = warp % 4).
// Persistent over tiles (tile = image x phase x BHxBW pixel patch); STAGES-deep smem ring of (A tile, W tile)
// K-blocks guarded by full/empty mbarriers; two TMEM accumulator stages so the epilogue of tile i overlaps
// the MMAs of tile i+1.
#ifdef FCB_EXPERIMENT
#include <algorithm>
#include <vector>

#include "fcb_epilogue.cuh"
#include "fcb_sm100.cuh"
#include "fcb_umma_common.h"

namespace fcb {

using namespace sm100;

constexpr int KCH = 128;      // bytes of K per pipeline stage = one 128B swizzle row
constexpr int TILE_M = 128;   // output pixels per tile (UMMA M)
constexpr int MAX_TAPS = 32;
constexpr int NUM_THREADS = 192;

struct Tap {
  int8_t offx, offy;  // input offset of the tap, in units of the (parity-split) input grid
  int8_t parx, pary;  // parity plane (stride 2 only)
  int32_t wtap;       // ky*KX + kx : selects the K range [wtap*C, (wtap+1)*C) of the weights
};
struct Phase {
  int ntaps, px, py, pad_;
  Tap taps[MAX_TAPS];
};
struct UmmaParams {
  uint8_t* out;
  EpiParams epi;
  int C, N, OFM, cchunks;
  int stride2, deconv, nphases;
  int BW, BH, tiles_x, tiles_y;  // per phase
  int PX, PY;                    // extent of the tile grid's pixel domain (output pixels per phase)
  int OX, OY, out_x, out_y, out_word_bytes;
  int n_images, stages, acc_stride, tmem_cols;
  unsigned long long out_img_bytes;
  uint32_t idesc;
  Phase phases[4];
};

struct UmmaV1 {
  Geom g;
  CUtensorMap tmB;
  UmmaParams p;
  size_t smem = 0;
  int num_sms = 148;
};

// ------------------------------------------------------------------------------------------
template <bool STRIDE2>
__global__ void __launch_bounds__(NUM_THREADS, 1)
umma_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = TILE_M * KCH, b_bytes = p.N * KCH, stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* tfull = bars + 2 * p.stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tiles_per_img = (long long)p.nphases * p.tiles_x * p.tiles_y;
  const long long total_tiles = tiles_per_img * p.n_images;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int img = (int)(t / tiles_per_img);
        int r = (int)(t % tiles_per_img);
        const int ph = r / (p.tiles_x * p.tiles_y);
        r %= p.tiles_x * p.tiles_y;
        const int ty = r / p.tiles_x, tx = r % p.tiles_x;
        const Phase& P = p.phases[ph];
        for (int tp = 0; tp < P.ntaps; tp++) {
          const Tap tap = P.taps[tp];
          for (int cc = 0; cc < p.cchunks; cc++, it++) {
            const int s = it % p.stages;
            mbar_wait(&empty[s], ((it / p.stages) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
            uint8_t* a_dst = smem + (size_t)s * stage_bytes;
            if (STRIDE2)
              tma_load_5d(a_dst, &tmA, &full[s], tap.parx * p.C + cc * KCH, tx * p.BW + tap.offx, tap.pary, ty * p.BH + tap.offy, img);
            else
              tma_load_4d(a_dst, &tmA, &full[s], cc * KCH, tx * p.BW + tap.offx, ty * p.BH + tap.offy, img);
            tma_load_2d(a_dst + a_bytes, &tmB, &full[s], tap.wtap * p.C + cc * KCH, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, tile_it = 0;
      for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x, tile_it++) {
        const int ph = (int)((t % tiles_per_img) / (p.tiles_x * p.tiles_y));
        const int nkb = p.phases[ph].ntaps * p.cchunks;
        const int acc = tile_it & 1;
        mbar_wait(&tempty[acc], ((tile_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % p.stages;
          mbar_wait(&full[s], (it / p.stages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t adesc = make_smem_desc(a_addr, 128), bdesc = make_smem_desc(a_addr + a_bytes, 128);
#pragma unroll
          for (int k = 0; k < KCH / 32; k++)
            umma_i8(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty[s]);  // frees the smem slot when these MMAs retire
        }
        umma_commit(&tfull[acc]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;    // tile row = pixel index within the patch
    const int pw = row % p.BW, phh = row / p.BW;
    const int pk = p.epi.pool >= 2 ? p.epi.pool : 1;
    uint32_t tile_it = 0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x, tile_it++) {
      const int img = (int)(t / tiles_per_img);
      int r = (int)(t % tiles_per_img);
      const int ph = r / (p.tiles_x * p.tiles_y);
      r %= p.tiles_x * p.tiles_y;
      const int ty = r / p.tiles_x, tx = r % p.tiles_x;
      const int acc = tile_it & 1;
      // pixel of this thread in the pre-pool output map
      const int gx = tx * p.BW + pw, gy = ty * p.BH + phh;
      const bool inside = gx < p.PX && gy < p.PY;
      const int ox = p.deconv ? 2 * gx + p.phases[ph].px : gx;
      const int oy = p.deconv ? 2 * gy + p.phases[ph].py : gy;
      uint8_t* word = p.out + (size_t)img * p.out_img_bytes + ((size_t)(oy / pk) * p.out_x + (ox / pk)) * p.out_word_bytes;
      const bool writer = inside && (pk == 1 || ((ox % pk) == 0 && (oy % pk) == 0));

      mbar_wait(&tfull[acc], (tile_it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (p.epi.act_kind == FCB_ACT_BIAS_RELU && p.epi.out_bits == 8 && p.epi.acc_bits == 8 && pk == 1) {
          // fast path of the reference network: ((acc mod 256) + bias) mod 256, MSB -> 0
          uint32_t w4[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            uint32_t pack = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
              const int ch = c0 + j * 4 + b;
              uint32_t rr = (v[j * 4 + b] + (uint32_t)(int32_t)p.epi.bias[ch]) & 0xFFu;
              rr = (rr & 0x80u) ? 0u : rr;
              pack |= rr << (8 * b);
            }
            w4[j] = pack;
          }
          if (inside && c0 < p.OFM) {
            uint4* dst = reinterpret_cast<uint4*>(word + c0);
            dst[0] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            dst[1] = make_uint4(w4[4], w4[5], w4[6], w4[7]);
          }
        } else {
          // general path: activation per lane, optional max pool across the patch, bit-packed store
          uint32_t packed[32];  // enough for 32 lanes x 32 bits
#pragma unroll
          for (int j = 0; j < 32; j++) packed[j] = 0;
          const int ob = p.epi.out_bits;
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int ch = c0 + j;
            uint32_t a = (ch < p.OFM) ? activate(p.epi, ch, (int32_t)v[j]) : 0u;
            if (pk == 2) {  // BW <= 16: the 2x2 window lives in lanes {l, l^1, l^BW, l^BW^1}
              a = max(a, __shfl_xor_sync(0xffffffffu, a, 1));
              a = max(a, __shfl_xor_sync(0xffffffffu, a, p.BW));
            }
            const int bit = j * ob;
            if (ob == 32) packed[j] = a;
            else packed[bit >> 5] |= a << (bit & 31);
          }
          if (writer) {
            const int nbytes = 4 * ob;  // 32 lanes x ob bits
            uint8_t* dst = word + ((size_t)c0 * ob >> 3);
            const int valid_bytes = min(nbytes, (int)((((size_t)p.OFM - c0) * ob + 7) >> 3));
            if (valid_bytes == nbytes) {
              for (int w = 0; w < ob; w++) reinterpret_cast<uint32_t*>(dst)[w] = packed[w];
            } else {
              for (int b = 0; b < valid_bytes; b++) dst[b] = (uint8_t)(packed[b >> 2] >> (8 * (b & 3)));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------
static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// v1 (per-tap TMA) needs whole 128-byte channel chunks and 32-channel epilogue groups
int umma_v1_eligible(const Geom& g) { return g.DX == 1 && g.DY == 1 && g.C % KCH == 0 && g.OFM % 32 == 0 && g.OFM >= 32 && g.OFM <= 256 && g.KX * g.KY <= MAX_TAPS; }

int umma_v1_create(const Geom& g, int8_t* d_w, const EpiParams& epi, int num_sms, UmmaV1** out) {
  *out = nullptr;
  if (!umma_v1_eligible(g)) return FCB_ERR_UNSUPPORTED;
  UmmaV1* P = new UmmaV1();
  P->g = g;
  P->num_sms = num_sms;
  UmmaParams& p = P->p;
  memset(&p, 0, sizeof(p));
  p.epi = epi;
  p.C = g.C; p.OFM = g.OFM; p.N = g.OFM; p.cchunks = (g.C + KCH - 1) / KCH;
  p.deconv = g.kind == FCB_KIND_DECONV522;
  p.stride2 = (!p.deconv && g.SX == 2);
  p.OX = g.OX; p.OY = g.OY; p.out_x = g.out_x; p.out_y = g.out_y; p.out_word_bytes = (int)g.out_word_bytes;
  p.out_img_bytes = g.out_img_bytes;
  p.PX = p.deconv ? g.IX : g.OX;  // per-phase pixel domain
  p.PY = p.deconv ? g.IY : g.OY;
  // tile shape: BW x BH = 128 pixels; prefer wide tiles, keep waste small; pooled layers need BW <= 16
  {
    int best_bw = 8;
    double best = 1e30;
    const int max_bw = g.pool == 2 ? 16 : 128;
    for (int bw = 8; bw <= max_bw; bw *= 2) {
      const int bh = TILE_M / bw;
      const double tiles = (double)((p.PX + bw - 1) / bw) * ((p.PY + bh - 1) / bh);
      const double cost = tiles * (1.0 + 0.02 * bh);  // mild preference for wide boxes (longer TMA rows)
      if (cost < best) { best = cost; best_bw = bw; }
    }
    p.BW = best_bw; p.BH = TILE_M / best_bw;
  }
  p.tiles_x = (p.PX + p.BW - 1) / p.BW;
  p.tiles_y = (p.PY + p.BH - 1) / p.BH;
  if (!p.deconv) {
    p.nphases = 1;
    Phase& ph = p.phases[0];
    ph.ntaps = 0; ph.px = ph.py = 0;
    for (int ky = 0; ky < g.KY; ky++)
      for (int kx = 0; kx < g.KX; kx++) {
        Tap t;
        const int dx = kx - g.PAD, dy = ky - g.PAD;
        t.offx = (int8_t)floordiv(dx, g.SX); t.parx = (int8_t)(dx - floordiv(dx, g.SX) * g.SX);
        t.offy = (int8_t)floordiv(dy, g.SY); t.pary = (int8_t)(dy - floordiv(dy, g.SY) * g.SY);
        t.wtap = ky * g.KX + kx;
        ph.taps[ph.ntaps++] = t;
      }
  } else {
    p.nphases = 4;
    for (int py = 0; py < 2; py++)
      for (int px = 0; px < 2; px++) {
        Phase& ph = p.phases[py * 2 + px];
        ph.ntaps = 0; ph.px = px; ph.py = py;
        for (int ky = 0; ky < 5; ky++)
          for (int kx = 0; kx < 5; kx++) {
            if (((px + kx) & 1) || ((py + ky) & 1)) continue;  // structural zero (A.6)
            Tap t;
            t.offx = (int8_t)((px + kx - 2) / 2); t.offy = (int8_t)((py + ky - 2) / 2);
            t.parx = t.pary = 0;
            t.wtap = ky * 5 + kx;
            ph.taps[ph.ntaps++] = t;
          }
      }
  }
  p.idesc = make_idesc_i8(TILE_M, p.N, g.in_signed, 1);
  int pow2 = 32;
  while (pow2 < p.N) pow2 *= 2;
  p.acc_stride = pow2;
  p.tmem_cols = 2 * pow2;
  const int stage_bytes = TILE_M * KCH + p.N * KCH;
  p.stages = std::min(8, (int)((200 * 1024) / stage_bytes));
  P->smem = (size_t)p.stages * stage_bytes + 1024 /*align*/ + 256 /*barriers*/;
  const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)p.N};
  const uint64_t strides[1] = {(uint64_t)g.K};
  const uint32_t box[2] = {(uint32_t)KCH, (uint32_t)p.N};
  int rc = umma_encode_map(&P->tmB, d_w, 2, dims, strides, box);
  if (rc) { delete P; return rc; }
  cudaError_t e = cudaFuncSetAttribute(umma_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) { delete P; set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FCB_ERR_CUDA; }
  *out = P;
  return FCB_OK;
}

void umma_v1_destroy(UmmaV1* P) { delete P; }

int umma_v1_run(UmmaV1* P, const void* d_in, void* d_out, int n_images, cudaStream_t st) {
  const Geom& g = P->g;
  UmmaParams p = P->p;
  p.out = (uint8_t*)d_out;
  p.n_images = n_images;
  CUtensorMap tmA;
  int rc;
  const uint64_t C = g.C, X = g.IX, Y = g.IY;
  if (p.stride2) {
    // [n][y/2][y%2][x/2][(x%2)*C + c] view of the dense NHWC stream
    const uint64_t dims[5] = {2 * C, X / 2, 2, Y / 2, (uint64_t)n_images};
    const uint64_t strides[4] = {2 * C, X * C, 2 * X * C, X * Y * C};
    const uint32_t box[5] = {(uint32_t)KCH, (uint32_t)p.BW, 1, (uint32_t)p.BH, 1};
    rc = umma_encode_map(&tmA, const_cast<void*>(d_in), 5, dims, strides, box);
  } else {
    const uint64_t dims[4] = {C, X, Y, (uint64_t)n_images};
    const uint64_t strides[3] = {C, X * C, X * Y * C};
    const uint32_t box[4] = {(uint32_t)KCH, (uint32_t)p.BW, (uint32_t)p.BH, 1};
    rc = umma_encode_map(&tmA, const_cast<void*>(d_in), 4, dims, strides, box);
  }
  if (rc) return rc;
  const long long total = (long long)p.nphases * p.tiles_x * p.tiles_y * n_images;
  const int grid = (int)std::min<long long>(total, P->num_sms);
  if (p.stride2)
    umma_conv_kernel<true><<<grid, NUM_THREADS, P->smem, st>>>(tmA, P->tmB, p);
  else
    umma_conv_kernel<false><<<grid, NUM_THREADS, P->smem, st>>>(tmA, P->tmB, p);
  FCB_CUDA_OK(cudaGetLastError());
  return FCB_OK;
}

}  // namespace fcb
#endif  // FCB_EXPERIMENT
