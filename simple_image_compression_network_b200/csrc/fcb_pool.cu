// fcb_pool.cu -- one layer chain replicated on several GPUs of the box behind ONE handle: the in-library form of SURVEY.md 8(e).
//
// The reference's top function takes the whole batch in one call -- eight_layers_net(in, out, numReps), conv_nonsquare_top.cpp:295 --
// and every image (numReps index) is an independent application of the layers (the sliding-window buffers reset per image,
// slidingwindow.h:1320,1351).  fcb_pool_run keeps that call shape and splits numReps into contiguous image ranges, one per replica,
// each served by its own host thread, device, streams and staging slots (fcb_net_run / fcb_layer_run underneath).  Weights are
// replicated (<= 0.4 MB per layer); there is no data-path exchange between devices, hence no collective.
#include <string>
#include <thread>
#include <vector>

#include "fcb_internal.h"

using namespace fcb;

struct fcb_pool {
  struct Replica {
    int device = 0;
    std::vector<fcb_layer*> layers;
    fcb_net* net = nullptr;  // chains of more than one layer
    int rc = FCB_OK;
    std::string err;
  };
  std::vector<Replica> reps;
  size_t in_img_bytes = 0, out_img_bytes = 0;
};

extern "C" {

int fcb_shard_range(uint32_t numReps, uint32_t rank, uint32_t world, uint32_t* begin, uint32_t* end) {
  if (!world || rank >= world || !begin || !end) { set_error("fcb_shard_range: bad arguments"); return FCB_ERR_INVALID_ARG; }
  const uint32_t base = numReps / world, rem = numReps % world;  // the remainder goes to the low ranks, one image each
  *begin = rank * base + (rank < rem ? rank : rem);
  *end = *begin + base + (rank < rem ? 1u : 0u);
  return FCB_OK;
}

void fcb_pool_destroy(fcb_pool* P) {
  if (!P) return;
  for (auto& r : P->reps) {
    if (r.net) fcb_net_destroy(r.net);
    for (fcb_layer* l : r.layers) fcb_layer_destroy(l);
  }
  delete P;
}

int fcb_pool_create(const fcb_layer_desc* descs, const void* const* weights, const void* const* thresholds, const void* const* biases,
                    uint32_t n_layers, const int* devices, uint32_t n_devices, fcb_pool** out) {
  if (!out) { set_error("out is NULL"); return FCB_ERR_INVALID_ARG; }
  *out = nullptr;
  if (!descs || !weights || !n_layers) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  std::vector<int> devs;
  if (devices && n_devices) devs.assign(devices, devices + n_devices);
  else {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    for (int i = 0; i < n; i++) {
      int major = 0;
      if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) devs.push_back(i);
    }
  }
  if (devs.empty()) { set_error("no usable sm_100 device (this library has no CPU path)"); return FCB_ERR_CUDA; }
  // the array's element size is the caller's struct_size (a client built against ABI 0.1 passes shorter structs)
  std::vector<fcb_layer_desc> dn(n_layers);
  for (uint32_t i = 0; i < n_layers; i++) {
    int rc = normalize_desc((const fcb_layer_desc*)((const uint8_t*)descs + (size_t)i * descs->struct_size), &dn[i]);
    if (rc) return rc;
  }
  descs = dn.data();
  fcb_pool* P = new fcb_pool();
  P->reps.resize(devs.size());
  for (size_t r = 0; r < devs.size(); r++) {
    fcb_pool::Replica& R = P->reps[r];
    R.device = devs[r];
    for (uint32_t i = 0; i < n_layers; i++) {
      fcb_layer* L = nullptr;
      int rc = fcb_layer_create(&descs[i], weights[i], thresholds ? thresholds[i] : nullptr, biases ? biases[i] : nullptr, R.device, &L);
      if (rc) { fcb_pool_destroy(P); return rc; }
      R.layers.push_back(L);
    }
    if (n_layers > 1) {
      int rc = fcb_net_create(R.layers.data(), n_layers, &R.net);
      if (rc) { fcb_pool_destroy(P); return rc; }
    }
  }
  size_t ib = 0, ob = 0;
  fcb_layer_query(&descs[0], &ib, nullptr, nullptr, nullptr, nullptr);
  fcb_layer_query(&descs[n_layers - 1], nullptr, &ob, nullptr, nullptr, nullptr);
  P->in_img_bytes = ib; P->out_img_bytes = ob;
  *out = P;
  return FCB_OK;
}

uint32_t fcb_pool_replicas(const fcb_pool* P) { return P ? (uint32_t)P->reps.size() : 0; }

int fcb_pool_run(fcb_pool* P, const void* in_words, void* out_words, uint32_t numReps) {
  if (!P || !in_words || !out_words) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  if (!numReps) return FCB_OK;
  const uint32_t world = (uint32_t)P->reps.size();
  auto work = [&](uint32_t r) {
    fcb_pool::Replica& R = P->reps[r];
    uint32_t b = 0, e = 0;
    fcb_shard_range(numReps, r, world, &b, &e);
    R.rc = FCB_OK;
    if (e == b) return;
    const uint8_t* src = (const uint8_t*)in_words + (size_t)b * P->in_img_bytes;
    uint8_t* dst = (uint8_t*)out_words + (size_t)b * P->out_img_bytes;
    R.rc = R.net ? fcb_net_run(R.net, src, dst, e - b) : fcb_layer_run(R.layers[0], src, dst, e - b);
    if (R.rc) R.err = fcb_last_error();  // (the message is thread-local: carry it to the caller's thread)
  };
  std::vector<std::thread> th;
  for (uint32_t r = 1; r < world; r++) th.emplace_back(work, r);
  work(0);
  for (auto& t : th) t.join();
  for (uint32_t r = 0; r < world; r++)
    if (P->reps[r].rc) {
      set_error("replica %u (device %d): %s", r, P->reps[r].device, P->reps[r].err.c_str());
      return P->reps[r].rc;
    }
  return FCB_OK;
}

// Page-locked host memory visible to every device: what the staging copies of the host-buffer calls run at full PCIe rate from
// (pageable memory is bounced through a driver buffer).  Plain malloc'ed buffers are accepted everywhere; these are faster.
int fcb_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) { set_error("NULL argument"); return FCB_ERR_INVALID_ARG; }
  *ptr = nullptr;
  cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? FCB_ERR_NOMEM : FCB_ERR_CUDA;
  }
  return FCB_OK;
}

void fcb_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
