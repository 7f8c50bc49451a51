// fcb_direct.cu -- direct (CUDA-core) convolution engines: "imad" and "xnor_popc".
//
// imad      : the universal integer path (any lane width <= 16 bits, +-1 / xnor weights, any activation).
//             north_star: "IMAD for wider types".  Replaces, in one kernel,
//             FMPadding_nonsquare (streamtools.h:361-406), the zero-insertion of deconv522
//             (conv_nonsquare_top.cpp:109-156), ConvolutionInputGenerator_NonSquare
//             (slidingwindow.h:1254-1353), Matrix_Vector_Activate_Batch (mvau.hpp:87-179) and the
//             activation / bias-ReLU / pool stages.
// xnor_popc : BinaryWeights + Recast<XnorMul> (weights.hpp:66-98, interpret.hpp:57-73) on bit-packed
//             lanes: acc = sum over 32-bit words of popc(~(a ^ w)).
//
// Work decomposition (both): one CTA = 16x8 pre-pool output pixels x 64 output channels of one
// image; a warp owns a 4x4 pixel block, a lane owns channels {l, l+32}; the input patch of the tile
// is staged once in shared memory (zero padding / zero insertion resolved while staging), weights
// are read through L1 as [k][channel] rows so a warp's 32 lanes load 32 consecutive values.
#include "fcb_epilogue.cuh"

namespace fcb {

constexpr int TX = 16, TY = 8, PXB = 4;  // tile and per-warp pixel block
constexpr int CH_PER_CTA = 64;

// Maps a coordinate of the virtual padded frame to the input image; false = structural zero.
__device__ __forceinline__ bool map_coord(int v, int pad, int deconv, int extent, int* src) {
  if (!deconv) {
    const int s = v - pad;
    *src = s;
    return s >= 0 && s < extent;
  }
  const int s = v - 2;  // deconv522: Z(2i,2j) = a(i,j), frame padded by 2
  *src = s >> 1;
  return s >= 0 && !(s & 1) && (s >> 1) < extent;
}

// activation, pool, store of one warp's 4x4 pixel block x 2 channels per lane (shared by both direct kernels)
__device__ __forceinline__ void direct_epilogue(const DirectParams& p, const int32_t (&acc)[PXB * PXB][2], int img, int ch0, int ox0, int oy0, int wx,
                                                int wy, int lane) {
  // ---- activation, pool, store -----------------------------------------------------------
  const int pk = p.epi.pool >= 2 ? p.epi.pool : 1;
  uint8_t* out = p.out + (size_t)img * p.out_img_bytes;
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int ch = ch0 + 32 * j;
    const bool chv = ch < p.OFM;
    uint32_t val[PXB * PXB];
#pragma unroll
    for (int i = 0; i < PXB * PXB; i++) val[i] = chv ? activate(p.epi, ch, acc[i][j]) : 0u;
    if (p.epi.out_bits == 1 && pk == 1) {
      // 1-bit lanes: one ballot per pixel; lane i keeps pixel i of the 4x4 block, then one store instruction for all 16
      uint32_t mine = 0;
#pragma unroll
      for (int i = 0; i < PXB * PXB; i++) {
        const uint32_t bits = __ballot_sync(0xffffffffu, chv && (val[i] & 1u));
        if (lane == i) mine = bits;
      }
      const int oy = oy0 + wy + (lane >> 2), ox = ox0 + wx + (lane & 3), c0 = ch - lane;  // c0: first channel of this warp
      if (lane < PXB * PXB && oy < p.OY && ox < p.OX && c0 < p.OFM) {
        uint8_t* dst = out + ((size_t)oy * p.out_x + ox) * p.out_word_bytes + (c0 >> 3);
        if (c0 + 32 <= p.OFM && p.out_word_bytes >= 4) *reinterpret_cast<uint32_t*>(dst) = mine;
        else
          for (int b = 0; b < 4; b++)
            if (c0 + 8 * b < p.OFM) dst[b] = (uint8_t)(mine >> (8 * b));
      }
      continue;
    }
    for (int by = 0; by < PXB; by += pk)
      for (int bx = 0; bx < PXB; bx += pk) {
        const int oy = oy0 + wy + by, ox = ox0 + wx + bx;  // pre-pool pixel (warp-uniform)
        if (oy >= p.OY || ox >= p.OX) continue;
        uint32_t m = 0;
        for (int dy = 0; dy < pk; dy++)
          for (int dx = 0; dx < pk; dx++) m = max(m, val[(by + dy) * PXB + bx + dx]);
        uint8_t* word = out + ((size_t)(oy / pk) * p.out_x + (ox / pk)) * p.out_word_bytes;
        store_lane(word, ch, chv, m, p.epi.out_bits);
      }
  }
}

template <int ENGINE>
__global__ void __launch_bounds__(256) direct_conv_kernel(const DirectParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, tx = tile % p.tiles_x, ty = tile / p.tiles_x;
  const int img = blockIdx.z;
  const int ch0 = blockIdx.y * CH_PER_CTA + lane;
  const int ox0 = tx * TX, oy0 = ty * TY;
  const int wx = (warp & 3) * PXB, wy = (warp >> 2) * PXB;  // this warp's 4x4 pixel block in the tile
  const uint8_t* in = p.in + (size_t)img * p.in_img_bytes;

  int32_t acc[PXB * PXB][2];
#pragma unroll
  for (int i = 0; i < PXB * PXB; i++) acc[i][0] = acc[i][1] = 0;

  const int vx0 = ox0 * p.SXe, vy0 = oy0 * p.SYe;  // tile origin in the padded frame
  // channel units: imad = lanes, xnor = 32-bit words of packed lanes
  const int CU = (ENGINE == ENG_XNOR) ? (p.C >> 5) : p.C;

  for (int c0 = 0; c0 < CU; c0 += p.CC) {
    const int cc = min(p.CC, CU - c0);
    // ---- stage the patch ------------------------------------------------------------
    const int total = p.patch_h * p.patch_w * cc;
    for (int idx = tid; idx < total; idx += blockDim.x) {
      const int c = idx % cc, pix = idx / cc;
      const int px = pix % p.patch_w, py = pix / p.patch_w;
      int sx, sy;
      const bool okx = map_coord(vx0 + px, p.PAD, p.deconv, p.IX, &sx);
      const bool oky = map_coord(vy0 + py, p.PADY, p.deconv, p.IY, &sy);  // (FMPadding_nonsquare may pad left / up differently)
      if (ENGINE == ENG_XNOR) {
        uint32_t v = 0;
        if (okx && oky) v = reinterpret_cast<const uint32_t*>(in + ((size_t)sy * p.IX + sx) * p.in_word_bytes)[c0 + c];
        reinterpret_cast<uint32_t*>(smem_raw)[idx] = v;
      } else {
        int32_t v = 0;
        if (okx && oky) v = load_lane_any(in + ((size_t)sy * p.IX + sx) * p.in_word_bytes, c0 + c, p.in_bits, p.in_signed);
        reinterpret_cast<int32_t*>(smem_raw)[idx] = v;
      }
    }
    __syncthreads();
    // ---- multiply-accumulate ----------------------------------------------------------
    if (ENGINE == ENG_XNOR && (cc & 1) == 0) {
      // channel words in pairs: one 8-byte patch load feeds two popcounts per output, the two counts and the accumulator meet
      // in one 3-input add, and the weights of the next (tap, pair) are fetched while this one is counted
      const uint32_t* wt = reinterpret_cast<const uint32_t*>(p.wt);
      const uint2* patch2 = reinterpret_cast<const uint2*>(smem_raw);
      const int cc2 = cc >> 1, nit = p.KY * p.KX * cc2;
      int ky = 0, kx = 0, c2 = 0;
      auto wrow_of = [&](int ky_, int kx_, int c2_) { return (size_t)((ky_ * p.KX + kx_) * CU + c0 + 2 * c2_) * p.OFMp + ch0; };
      uint32_t wn[4];
      {
        const size_t r = wrow_of(0, 0, 0);
        wn[0] = __ldg(wt + r); wn[1] = __ldg(wt + r + 32); wn[2] = __ldg(wt + r + p.OFMp); wn[3] = __ldg(wt + r + p.OFMp + 32);
      }
      for (int it = 0; it < nit; it++) {
        const uint32_t w00 = wn[0], w01 = wn[1], w10 = wn[2], w11 = wn[3];
        const int poff = (ky * p.DY * p.patch_w + kx * p.DX) * cc2 + c2;
        if (++c2 == cc2) { c2 = 0; if (++kx == p.KX) { kx = 0; ++ky; } }
        if (it + 1 < nit) {
          const size_t r = wrow_of(ky, kx, c2);
          wn[0] = __ldg(wt + r); wn[1] = __ldg(wt + r + 32); wn[2] = __ldg(wt + r + p.OFMp); wn[3] = __ldg(wt + r + p.OFMp + 32);
        }
#pragma unroll
        for (int i = 0; i < PXB * PXB; i++) {
          const int ly = wy + (i >> 2), lx = wx + (i & 3);
          const uint2 a = patch2[(ly * p.SYe * p.patch_w + lx * p.SXe) * cc2 + poff];
          acc[i][0] += __popc(~(a.x ^ w00)) + __popc(~(a.y ^ w10));
          acc[i][1] += __popc(~(a.x ^ w01)) + __popc(~(a.y ^ w11));
        }
      }
    } else
    for (int ky = 0; ky < p.KY; ky++)
      for (int kx = 0; kx < p.KX; kx++) {
        const int kbase = (ky * p.KX + kx) * CU + c0;
        for (int c = 0; c < cc; c++) {
          const size_t wrow = (size_t)(kbase + c) * p.OFMp + ch0;
          if (ENGINE == ENG_XNOR) {
            const uint32_t* wt = reinterpret_cast<const uint32_t*>(p.wt);
            const uint32_t w0 = __ldg(wt + wrow), w1 = __ldg(wt + wrow + 32);
            const uint32_t* patch = reinterpret_cast<const uint32_t*>(smem_raw);
#pragma unroll
            for (int i = 0; i < PXB * PXB; i++) {
              const int ly = wy + (i >> 2), lx = wx + (i & 3);
              const uint32_t a = patch[((ly * p.SYe + ky * p.DY) * p.patch_w + (lx * p.SXe + kx * p.DX)) * cc + c];
              acc[i][0] += __popc(~(a ^ w0));
              acc[i][1] += __popc(~(a ^ w1));
            }
          } else {
            const int16_t* wt = reinterpret_cast<const int16_t*>(p.wt);
            const int32_t w0 = __ldg(wt + wrow), w1 = __ldg(wt + wrow + 32);
            const int32_t* patch = reinterpret_cast<const int32_t*>(smem_raw);
#pragma unroll
            for (int i = 0; i < PXB * PXB; i++) {
              const int ly = wy + (i >> 2), lx = wx + (i & 3);
              const int32_t a = patch[((ly * p.SYe + ky * p.DY) * p.patch_w + (lx * p.SXe + kx * p.DX)) * cc + c];
              if (p.mul_kind == FCB_W_BINARY_XNOR) {
                acc[i][0] += (a == w0);
                acc[i][1] += (a == w1);
              } else {
                acc[i][0] += a * w0;
                acc[i][1] += a * w1;
              }
            }
          }
        }
      }
    __syncthreads();
  }
  direct_epilogue(p, acc, img, ch0, ox0, oy0, wx, wy, lane);
}

// ------------------------------------------------------------------------------------------------------------------------
// imad_conv_kernel -- the IMAD engine proper (FixedPoint / +-1 weights; north_star: "IMAD for wider types").
// Same decomposition as above (CTA = 16x8 pixels x 64 channels, warp = 4x4 pixels, lane = channels {l, l+32}), with the inner loop
// built around the shared-memory pipe instead of L1:
//   * the weights of the channel chunk are staged in shared memory as packed pairs, word [tap][c][lane] = w(ch l) | w(ch l+32) << 16:
//     one conflict-free LDS per (tap, channel) instead of two L1 loads whose latency nothing hid;
//   * the patch is stored with a channel pitch that is a multiple of 4, so one LDS.128 brings 4 channels of a pixel: per 4 channels
//     a lane issues 16 + 4 shared loads for 128 IMADs (16 + 2 per 32 before);
//   * deconv522: taps are walked by parity class and only the pixels whose (row + ky, column + kx) parities hit a non-zero sample
//     of the zero-inserted frame are unrolled: 4x fewer IMADs, the structural zeros are never multiplied.
template <bool DECONV, int PKY, int PKX>
__device__ __forceinline__ void imad_taps(const DirectParams& p, const int32_t* __restrict__ patch, const uint32_t* __restrict__ wsm, int cc4, int lane,
                                          int wx, int wy, int32_t (&acc)[PXB * PXB][2]) {
  const int rowstep = p.SYe * p.patch_w * cc4, colstep = p.SXe * cc4;
  for (int ky = DECONV ? PKY : 0; ky < p.KY; ky += DECONV ? 2 : 1)
    for (int kx = DECONV ? PKX : 0; kx < p.KX; kx += DECONV ? 2 : 1) {
      const uint32_t* wrow = wsm + (size_t)(ky * p.KX + kx) * cc4 * 32 + lane;
      const int32_t* prow = patch + ((size_t)(wy * p.SYe + ky * p.DY) * p.patch_w + wx * p.SXe + kx * p.DX) * cc4;
      for (int c = 0; c < cc4; c += 4) {
        int32_t w0[4], w1[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t wp = wrow[(c + j) * 32];
          w0[j] = (int32_t)(int16_t)(wp & 0xFFFFu);
          w1[j] = (int32_t)wp >> 16;
        }
#pragma unroll
        for (int i = 0; i < PXB * PXB; i++) {
          const int ly = i >> 2, lx = i & 3;
          if (DECONV && (((ly ^ PKY) & 1) || ((lx ^ PKX) & 1))) continue;  // compile-time: structural zero of the inserted frame
          const int4 a = *reinterpret_cast<const int4*>(prow + ly * rowstep + lx * colstep + c);
          acc[i][0] += a.x * w0[0] + a.y * w0[1] + a.z * w0[2] + a.w * w0[3];
          acc[i][1] += a.x * w1[0] + a.y * w1[1] + a.z * w1[2] + a.w * w1[3];
        }
      }
    }
}

__global__ void __launch_bounds__(256, 2) imad_conv_kernel(const DirectParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, tx = tile % p.tiles_x, ty = tile / p.tiles_x;
  const int img = blockIdx.z;
  const int chb = blockIdx.y * CH_PER_CTA, ch0 = chb + lane;
  const int ox0 = tx * TX, oy0 = ty * TY;
  const int wx = (warp & 3) * PXB, wy = (warp >> 2) * PXB;
  const uint8_t* in = p.in + (size_t)img * p.in_img_bytes;
  int32_t acc[PXB * PXB][2];
#pragma unroll
  for (int i = 0; i < PXB * PXB; i++) acc[i][0] = acc[i][1] = 0;
  const int vx0 = ox0 * p.SXe, vy0 = oy0 * p.SYe;
  int32_t* patch = reinterpret_cast<int32_t*>(smem_raw);
  const int16_t* wt = reinterpret_cast<const int16_t*>(p.wt);
  for (int c0 = 0; c0 < p.C; c0 += p.CC) {
    const int cc = min(p.CC, p.C - c0), cc4 = (cc + 3) & ~3;
    uint32_t* wsm = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)p.patch_h * p.patch_w * cc4;
    // ---- stage the patch: zero padding / zero insertion resolved here; channels cc..cc4 are zeros
    const int total = p.patch_h * p.patch_w * cc4;
    for (int idx = tid; idx < total; idx += blockDim.x) {
      const int c = idx % cc4, pix = idx / cc4;
      const int px = pix % p.patch_w, py = pix / p.patch_w;
      int sx, sy;
      const bool okx = map_coord(vx0 + px, p.PAD, p.deconv, p.IX, &sx);
      const bool oky = map_coord(vy0 + py, p.PADY, p.deconv, p.IY, &sy);
      int32_t v = 0;
      if (c < cc && okx && oky) v = load_lane_any(in + ((size_t)sy * p.IX + sx) * p.in_word_bytes, c0 + c, p.in_bits, p.in_signed);
      patch[idx] = v;
    }
    // ---- stage the weights of this channel chunk as packed pairs
    const int wtotal = p.KX * p.KY * cc4 * 32;
    for (int idx = tid; idx < wtotal; idx += blockDim.x) {
      const int l = idx & 31, c = (idx >> 5) % cc4, tap = (idx >> 5) / cc4;
      uint32_t wp = 0;
      if (c < cc) {
        const size_t row = (size_t)(tap * p.C + c0 + c) * p.OFMp + chb + l;
        wp = ((uint32_t)(uint16_t)__ldg(wt + row)) | ((uint32_t)(uint16_t)__ldg(wt + row + 32) << 16);
      }
      wsm[idx] = wp;
    }
    __syncthreads();
    if (p.deconv) {
      imad_taps<true, 0, 0>(p, patch, wsm, cc4, lane, wx, wy, acc);
      imad_taps<true, 0, 1>(p, patch, wsm, cc4, lane, wx, wy, acc);
      imad_taps<true, 1, 0>(p, patch, wsm, cc4, lane, wx, wy, acc);
      imad_taps<true, 1, 1>(p, patch, wsm, cc4, lane, wx, wy, acc);
    } else {
      imad_taps<false, 0, 0>(p, patch, wsm, cc4, lane, wx, wy, acc);
    }
    __syncthreads();
  }
  direct_epilogue(p, acc, img, ch0, ox0, oy0, wx, wy, lane);
}

size_t direct_smem_bytes(int engine, int patch_w, int patch_h, int cc) {
  (void)engine;
  return (size_t)patch_w * patch_h * cc * 4;
}
// imad_conv_kernel: patch (channel pitch rounded up to 4) + packed weight pairs of the chunk
size_t imad_smem_bytes(int patch_w, int patch_h, int taps, int cc) {
  const size_t cc4 = (size_t)((cc + 3) & ~3);
  return (size_t)patch_w * patch_h * cc4 * 4 + (size_t)taps * cc4 * 128;
}

int launch_direct(const DirectParams& p, int engine, int n_images, size_t smem_bytes, cudaStream_t st) {
  dim3 grid(p.tiles_x * p.tiles_y, (p.OFM + CH_PER_CTA - 1) / CH_PER_CTA, 1);
  // grid.z is limited to 65535 images per launch
  for (int n0 = 0; n0 < n_images; n0 += 65535) {
    DirectParams q = p;
    const int nb = n_images - n0 < 65535 ? n_images - n0 : 65535;
    q.in = p.in + (size_t)n0 * p.in_img_bytes;
    q.out = p.out + (size_t)n0 * p.out_img_bytes;
    grid.z = nb;
    if (engine == ENG_XNOR) {
      FCB_CUDA_OK(cudaFuncSetAttribute(direct_conv_kernel<ENG_XNOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      direct_conv_kernel<ENG_XNOR><<<grid, 256, smem_bytes, st>>>(q);
    } else if (p.mul_kind == FCB_W_BINARY_XNOR) {  // xnor layers the popcount engine cannot take (IFM_CH % 32 != 0): a == w per lane
      FCB_CUDA_OK(cudaFuncSetAttribute(direct_conv_kernel<ENG_IMAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      direct_conv_kernel<ENG_IMAD><<<grid, 256, smem_bytes, st>>>(q);
    } else {
      FCB_CUDA_OK(cudaFuncSetAttribute(imad_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      imad_conv_kernel<<<grid, 256, smem_bytes, st>>>(q);
    }
    FCB_CUDA_OK(cudaGetLastError());
  }
  return FCB_OK;
}

}  // namespace fcb
