// Library runtime: error slot, device gate (sm_100 only, no fallback), launch counter, TMA descriptor encoding
// through the driver entry point (so the .so links against nothing but the static CUDA runtime and still loads
// on a box without libcuda — the C-ABI export test runs there).
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "mla_internal.cuh"

namespace mla {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int g_sms = 0;
static int g_dev_rc = 1;  // 1 = not probed yet

int device_check() {
  if (g_dev_rc <= 0) return g_dev_rc;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(MLA_ERR_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return set_error(MLA_ERR_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (p.major != 10) {
    return set_error(MLA_ERR_DEVICE, "libmla_b200 is sm_100a-only; device %d is sm_%d%d (%s)", dev, p.major, p.minor,
                     p.name);
  }
  g_sms = p.multiProcessorCount;
  g_dev_rc = 0;
  return 0;
}

int num_sms() { return g_sms > 0 ? g_sms : 148; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// Descriptor cache: a CUtensorMap is a pure function of (pointer, rank, dims, strides, box) — it holds no reference to
// the memory's contents — and the training step asks for the same few hundred of them every step (the per-layer weight,
// activation and gradient buffers keep their addresses).  Direct-mapped, per host thread, so there is no lock on the
// launch path; a miss costs one driver encode (~1-2 us), a hit a 72-byte compare.
struct TmapKey {
  const void* ptr;
  uint64_t dims[3], strides[2];
  uint32_t box[3], rank;
};
struct TmapEntry {
  TmapKey key;
  CUtensorMap map;
  bool valid;
};
constexpr int kTmapCacheSize = 2048;
static std::atomic<int64_t> g_tmap_hits{0}, g_tmap_misses{0};

static int encode_nd_uncached(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                              const uint32_t* box);

static int encode_nd(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box) {
  static thread_local TmapEntry* cache = nullptr;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MLA_TMAP_CACHE");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return encode_nd_uncached(map, ptr, rank, dims, strides, box);
  if (!cache) cache = static_cast<TmapEntry*>(calloc(kTmapCacheSize, sizeof(TmapEntry)));
  if (!cache || rank > 3) return encode_nd_uncached(map, ptr, rank, dims, strides, box);
  TmapKey k;
  memset(&k, 0, sizeof(k));
  k.ptr = ptr;
  k.rank = uint32_t(rank);
  for (int i = 0; i < rank; ++i) k.dims[i] = dims[i], k.box[i] = box[i];
  for (int i = 0; i < rank - 1; ++i) k.strides[i] = strides[i];
  uint64_t h = reinterpret_cast<uint64_t>(ptr) >> 8;
  h = (h ^ (k.dims[0] * 0x9E3779B97F4A7C15ull) ^ (k.dims[1] * 0xC2B2AE3D27D4EB4Full) ^ (uint64_t(k.box[0]) << 17) ^
       (uint64_t(k.box[1]) << 29) ^ (k.strides[0] * 0x165667B19E3779F9ull)) *
      0xD6E8FEB86659FD93ull;
  TmapEntry& e = cache[(h >> 32) & (kTmapCacheSize - 1)];
  if (e.valid && memcmp(&e.key, &k, sizeof(k)) == 0) {
    *map = e.map;
    g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
    return MLA_OK;
  }
  if (int rc = encode_nd_uncached(map, ptr, rank, dims, strides, box)) return rc;
  e.key = k;
  e.map = *map;
  e.valid = true;
  g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
  return MLA_OK;
}

static int encode_nd_uncached(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                              const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(MLA_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t d[5];
  cuuint64_t s[5];
  cuuint32_t b[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) s[i] = strides[i];
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d, s, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(MLA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims {%llu,%llu} stride %llu box {%u,%u}",
                     int(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                     (unsigned long long)strides[0], box[0], box[1]);
  }
  return MLA_OK;
}

int encode_tmap_2d_bf16(CUtensorMap* map, const void* ptr, const uint64_t dims[2], const uint64_t strides[1],
                        const uint32_t box[2]) {
  return encode_nd(map, ptr, 2, dims, strides, box);
}
int encode_tmap_3d_bf16(CUtensorMap* map, const void* ptr, const uint64_t dims[3], const uint64_t strides[2],
                        const uint32_t box[3]) {
  return encode_nd(map, ptr, 3, dims, strides, box);
}

}  // namespace mla

extern "C" const char* mla_version(void) { return "mla_b200 0.1.0 (sm_100a)"; }
extern "C" const char* mla_last_error(void) { return mla::g_err; }
extern "C" int mla_device_check(void) { return mla::device_check(); }
extern "C" void mla_tmap_cache_stats(int64_t* hits, int64_t* misses) {
  if (hits) *hits = mla::g_tmap_hits.load(std::memory_order_relaxed);
  if (misses) *misses = mla::g_tmap_misses.load(std::memory_order_relaxed);
}
extern "C" int64_t mla_launch_count(void) { return mla::g_launches.load(std::memory_order_relaxed); }
