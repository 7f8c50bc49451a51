// tools/umma_probe.cu -- hardware probes for the building blocks of the umma_i8 engine (run on a B200):
//   P1  TMA SWIZZLE_128B tile layout in shared memory (chunk16 ^= row & 7)
//   P2  tcgen05.mma kind::i8 (u8 x s8 -> s32) with K-major SW128 smem descriptors, TMEM read-back
//   P3  A-operand descriptors whose start address is shifted by whole 128-byte rows (with / without the
//       descriptor's base_offset field) -- the premise of a smem-resident input patch shared by all filter taps
//   P4  5-D "parity view" tensor map with negative / out-of-range box coordinates (zero fill = FMPadding)
//   P5  4-D tensor map with elementStrides = 2 (alternative way to express the stride-2 window)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../simple_image_compression_network_b200/csrc/fcb_sm100.cuh"

using namespace fcb::sm100;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e = (x);                                                                        \
    if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_enc;
static bool make_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                     const uint32_t* estr) {
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = estr ? estr[i] : 1; }
  for (int i = 0; i + 1 < rank; i++) gs[i] = strides[i];
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("  cuTensorMapEncodeTiled rank %d failed: CUresult %d\n", rank, (int)r); return false; }
  return true;
}

constexpr int A_ROWS = 256, NVAR = 10;
struct Variant { int shift, bo; };
__constant__ Variant c_var[NVAR];

// ---- P1..P3 ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
probe_mma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, uint8_t* dumpA, int32_t* out, int a_signed) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // 256 x 128 B
  uint8_t* sB = smem + A_ROWS * 128;        // 128 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 128 * 128);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, (A_ROWS + 128) * 128);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < A_ROWS * 128; i += blockDim.x) dumpA[i] = sA[i];
  __syncthreads();
  const uint32_t idesc = make_idesc_i8(128, 128, a_signed, 1);
  for (int v = 0; v < NVAR; v++) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA) + c_var[v].shift * 128;
      const uint64_t adesc = make_smem_desc(a_addr, 128, c_var[v].bo), bdesc = make_smem_desc(smem_u32(sB), 128);
      for (int k = 0; k < 4; k++) umma_i8(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
      umma_commit(mbar);
    }
    mbar_wait(mbar, v & 1);
    tc_fence_after();
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; j++) out[((size_t)v * 128 + warp * 32 + lane) * 128 + c0 + j] = (int32_t)r[j];
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- P4 / P5: one box -> smem -> global dump ----------------------------------------------------
__global__ void probe_box(const __grid_constant__ CUtensorMap tm, uint8_t* dump, int rank, int c0, int c1, int c2, int c3, int c4, int bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) smem[i] = 0xEE;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, bytes);
    if (rank == 5) tma_load_5d(smem, &tm, bar, c0, c1, c2, c3, c4);
    else tma_load_4d(smem, &tm, bar, c0, c1, c2, c3);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) dump[i] = smem[i];
}

static inline int unswz(int o) {  // smem byte offset -> logical (row*128 + col) offset of a SW128 tile
  const int row = o / 128, chunk = (o % 128) / 16, b = o % 16;
  return row * 128 + ((chunk ^ (row & 7)) * 16) + b;
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  g_enc = (PFN_encodeTiled)p;
  if (!g_enc) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);

  // ================= P1-P3 =================
  std::vector<uint8_t> hA(A_ROWS * 128), hB(128 * 128);
  srand(1);
  for (auto& v : hA) v = rand() & 0xFF;
  for (auto& v : hB) v = rand() & 0xFF;
  uint8_t *dA, *dB, *dDump;
  int32_t* dOut;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dDump, hA.size()));
  CK(cudaMalloc(&dOut, (size_t)NVAR * 128 * 128 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {128, A_ROWS}, st[1] = {128};
    uint32_t box[2] = {128, A_ROWS};
    if (!make_map(&tmA, dA, 2, dims, st, box, nullptr)) return 2;
    uint64_t dimsB[2] = {128, 128};
    uint32_t boxB[2] = {128, 128};
    if (!make_map(&tmB, dB, 2, dimsB, st, boxB, nullptr)) return 2;
  }
  Variant var[NVAR] = {{0, 0}, {1, 0}, {1, 1}, {2, 0}, {2, 2}, {3, 3}, {5, 0}, {5, 5}, {8, 0}, {7, 7}};
  CK(cudaMemcpyToSymbol(c_var, var, sizeof(var)));
  const int smem = (A_ROWS + 128) * 128 + 1024 + 64;
  CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int a_signed = 0; a_signed < 2; a_signed++) {
    CK(cudaMemset(dOut, 0xCD, (size_t)NVAR * 128 * 128 * 4));
    probe_mma<<<1, 128, smem>>>(tmA, tmB, dDump, dOut, a_signed);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("P2 kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
    std::vector<uint8_t> dump(hA.size());
    std::vector<int32_t> out((size_t)NVAR * 128 * 128);
    CK(cudaMemcpy(dump.data(), dDump, dump.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
    if (!a_signed) {
      int bad = 0;
      for (int o = 0; o < A_ROWS * 128; o++) bad += dump[o] != hA[unswz(o)];
      printf("P1 TMA SW128 layout (phys chunk = logical chunk ^ (row&7)): %s (%d mismatching bytes)\n", bad ? "FAIL" : "PASS", bad);
    }
    for (int v = 0; v < NVAR; v++) {
      long bad = 0;
      int first = -1;
      for (int i = 0; i < 128; i++)
        for (int n = 0; n < 128; n++) {
          long acc = 0;
          for (int k = 0; k < 128; k++) {
            const int a = a_signed ? (int)(int8_t)hA[(var[v].shift + i) * 128 + k] : (int)hA[(var[v].shift + i) * 128 + k];
            acc += (long)a * (int)(int8_t)hB[n * 128 + k];
          }
          if ((int32_t)acc != out[((size_t)v * 128 + i) * 128 + n]) { if (first < 0) first = i * 128 + n; bad++; }
        }
      printf("%s a_%s shift=%d rows base_offset=%d : %s (%ld/16384 wrong%s", v == 0 ? "P2" : "P3", a_signed ? "s8" : "u8", var[v].shift,
             var[v].bo, bad ? "FAIL" : "PASS", bad, bad ? ", first at row " : "");
      if (bad) printf("%d col %d", first / 128, first % 128);
      printf(")\n");
    }
  }

  // ================= P4: 5-D parity view, OOB =================
  {
    const int C = 128, X = 16, Y = 8, N = 2;
    std::vector<uint8_t> img((size_t)N * Y * X * C);
    for (size_t i = 0; i < img.size(); i++) img[i] = 1 + (i * 2654435761u >> 24) % 255;
    uint8_t *dI, *dD;
    CK(cudaMalloc(&dI, img.size())); CK(cudaMalloc(&dD, 32768));
    CK(cudaMemcpy(dI, img.data(), img.size(), cudaMemcpyHostToDevice));
    CUtensorMap tm;
    const int BW = 8, BH = 4;
    uint64_t dims[5] = {2 * C, X / 2, 2, Y / 2, N}, st[4] = {2 * C, (uint64_t)X * C, 2ull * X * C, (uint64_t)X * Y * C};
    uint32_t box[5] = {128, BW, 1, BH, 1};
    CK(cudaFuncSetAttribute(probe_box, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024 + 64));
    if (make_map(&tm, dI, 5, dims, st, box, nullptr)) {
      struct { int parx, offx, pary, offy, n; } tests[] = {{0, -1, 0, -1, 1}, {1, 0, 1, 0, 0}, {1, 2, 0, 2, 1}, {0, 4, 1, 3, 0}};
      for (auto& t : tests) {
        probe_box<<<1, 128, 32768 + 1024 + 64>>>(tm, dD, 5, t.parx * C, t.offx, t.pary, t.offy, t.n, BW * BH * 128);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("P4 kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
        std::vector<uint8_t> d(BW * BH * 128);
        CK(cudaMemcpy(d.data(), dD, d.size(), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int o = 0; o < BW * BH * 128; o++) {
          const int l = unswz(o), row = l / 128, c = l % 128, w = row % BW, h = row / BW;
          const int x = 2 * (t.offx + w) + t.parx, y = 2 * (t.offy + h) + t.pary;
          const uint8_t exp = (x < 0 || x >= X || y < 0 || y >= Y) ? 0 : img[(((size_t)t.n * Y + y) * X + x) * C + c];
          bad += d[o] != exp;
        }
        printf("P4 5-D parity box parx=%d offx=%d pary=%d offy=%d n=%d : %s (%d wrong bytes)\n", t.parx, t.offx, t.pary, t.offy, t.n,
               bad ? "FAIL" : "PASS", bad);
      }
    }
    // ================= P5: elementStrides = 2 =================
    uint64_t dims4[4] = {(uint64_t)C, X, Y, N}, st4[3] = {(uint64_t)C, (uint64_t)X * C, (uint64_t)X * Y * C};
    uint32_t box4[4] = {128, 2 * BW, 2 * BH, 1}, es4[4] = {1, 2, 2, 1};
    if (make_map(&tm, dI, 4, dims4, st4, box4, es4)) {
      struct { int x0, y0, n; } tests[] = {{0, 0, 0}, {1, 1, 1}, {-2, -2, 0}, {-1, 3, 1}, {5, 2, 0}};
      for (auto& t : tests) {
        probe_box<<<1, 128, 32768 + 1024 + 64>>>(tm, dD, 4, 0, t.x0, t.y0, t.n, 0, BW * BH * 128);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("P5 kernel failed: %s (elementStrides unusable this way)\n", cudaGetErrorString(e)); break; }
        std::vector<uint8_t> d(BW * BH * 128);
        CK(cudaMemcpy(d.data(), dD, d.size(), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int o = 0; o < BW * BH * 128; o++) {
          const int l = unswz(o), row = l / 128, c = l % 128, w = row % BW, h = row / BW;
          const int x = t.x0 + 2 * w, y = t.y0 + 2 * h;
          const uint8_t exp = (x < 0 || x >= X || y < 0 || y >= Y) ? 0 : img[(((size_t)t.n * Y + y) * X + x) * C + c];
          bad += d[o] != exp;
        }
        printf("P5 elementStrides=2 box at x0=%d y0=%d n=%d : %s (%d wrong bytes)\n", t.x0, t.y0, t.n, bad ? "FAIL" : "PASS", bad);
      }
    }
  }
  printf("probe done\n");
  return 0;
}
