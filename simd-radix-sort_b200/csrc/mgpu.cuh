// mgpu.cuh -- multi-GPU sort: MSB range partition over the GPUs of one box.
// Included at the end of b200sort.cu (it uses that file's internals).
//
//   1. every rank histograms the top bits of its order-mapped keys        (top_hist_kernel)
//   2. ncclAllReduce(sum) of the histograms -> identical splitters on every rank (b200sort_mgpu_splitters)
//   3. ncclAllGather of the per-destination send counts -> receive offsets
//   4. local partition by destination rank = one scatter pass of onesweep_kernel in LUT mode
//   5. all-to-all-v: grouped ncclSend/ncclRecv per peer for the keys and every payload stream
//   6. local sort of the received range (sort_device)
// The reference has nothing here (single thread); NCCL is loaded lazily with dlopen so that the
// single-GPU library has no NCCL dependency.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace b200sort {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string why;
};

static NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
#define SYM(field, name)                                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));     \
    if (!api.field) { api.why = std::string("missing symbol ") + name; return; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    api.ok = true;
  });
  return api;
}

#define NCCL_TRY(expr)                                                                                   \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != ncclSuccess)                                                                               \
      return fail(B200SORT_ENCCL, "%s failed: %s", #expr, nccl_api().GetErrorString(_r));                 \
  } while (0)

constexpr int MGPU_MAX_BITS = 16;

// counts of the top `bits` bits of the ordered key; warp-aggregated global atomics
struct TopHistArgs {
  const unsigned char *keys;
  uint32_t stride;
  int64_t n;
  KeyOrder ko;
  int shift;            // ordered key >> shift = bin
  unsigned long long *hist;  // [2^bits], zeroed
};

template <int KB>
__global__ void __launch_bounds__(256) top_hist_kernel(TopHistArgs a) {
  const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = (a.n + 31) / 32 * 32;  // keep warps converged for the ballot
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gsz) {
    const bool valid = i < a.n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t bin = (uint32_t)(to_ordered<KB>(load_key<KB>(a.keys, i, a.stride), a.ko) >> a.shift);
      const unsigned peers = __match_any_sync(vmask, bin);
      if ((peers & lanemask_lt()) == 0) atomicAdd(&a.hist[bin], (unsigned long long)__popc(peers));
    }
  }
}

static cudaError_t launch_top_hist(int kb, const TopHistArgs &a, int sm_count, cudaStream_t st) {
  const int grid = (int)std::min<int64_t>((a.n + 255) / 256, (int64_t)sm_count * 16);
  switch (kb) {
    case 1: top_hist_kernel<1><<<grid, 256, 0, st>>>(a); break;
    case 2: top_hist_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 4: top_hist_kernel<4><<<grid, 256, 0, st>>>(a); break;
    default: top_hist_kernel<8><<<grid, 256, 0, st>>>(a); break;
  }
  g_launches++;
  return cudaGetLastError();
}

// Greedy splitters on bin boundaries: rank r ends at the first boundary where the running count
// reaches (r+1)/world of the total.  Identical on every rank because the input is the reduced histogram.
static void compute_splitters(const uint64_t *hist, int bits, int world, uint32_t *bounds) {
  const uint32_t nb = 1u << bits;
  uint64_t total = 0;
  for (uint32_t b = 0; b < nb; b++) total += hist[b];
  bounds[0] = 0;
  uint64_t run = 0;
  uint32_t b = 0;
  for (int r = 1; r < world; r++) {
    // smallest boundary whose prefix is >= ceil(total * r / world), but never before the previous one
    const uint64_t target = (uint64_t)(((unsigned __int128)total * (unsigned)r + (unsigned)world - 1) / (unsigned)world);
    while (b < nb && run < target) run += hist[b++];
    // choose the closer of the boundary just before and just after the target (bin granularity)
    if (b > bounds[r - 1] + 0u && b > 0) {
      const uint64_t over = run - target, under = target - (run - hist[b - 1]);
      if (under < over && b - 1 >= bounds[r - 1]) { run -= hist[b - 1]; b--; }
    }
    bounds[r] = b;
  }
  bounds[world] = nb;
}

}  // namespace b200sort

struct b200sort_comm {
  ncclComm_t comm = nullptr;
  int world = 0, rank = 0, dev = 0;
  unsigned long long *d_hist = nullptr;   // [2^16] local, then reduced in place
  unsigned long long *d_counts = nullptr; // [world] send counts, [world*world] gathered
  uint8_t *d_lut = nullptr;               // [2^16]
  uint64_t *d_bin_base = nullptr;         // [RADIX]
  b200sort::Plan *d_plan = nullptr;
};

extern "C" {

int b200sort_mgpu_unique_id(void *out) {
  using namespace b200sort;
  NcclApi &api = nccl_api();
  if (!api.ok) return fail(B200SORT_ENCCL, "NCCL unavailable: %s", api.why.c_str());
  if (!out) return fail(B200SORT_EINVAL, "out is NULL");
  static_assert(sizeof(ncclUniqueId) == B200SORT_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_TRY(api.GetUniqueId(&id));
  memcpy(out, &id, sizeof id);
  return 0;
}

int b200sort_mgpu_comm_create(b200sort_comm **out, int world_size, int rank, const void *id_bytes) {
  using namespace b200sort;
  NcclApi &api = nccl_api();
  if (!api.ok) return fail(B200SORT_ENCCL, "NCCL unavailable: %s", api.why.c_str());
  if (!out || !id_bytes || world_size < 1 || world_size > RADIX || rank < 0 || rank >= world_size)
    return fail(B200SORT_EINVAL, "bad communicator arguments");
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof id);
  b200sort_comm *c = new b200sort_comm();
  c->world = world_size;
  c->rank = rank;
  CUDA_TRY(cudaGetDevice(&c->dev));
  NCCL_TRY(api.CommInitRank(&c->comm, world_size, id, rank));
  CUDA_TRY(cudaMalloc(&c->d_hist, sizeof(unsigned long long) << MGPU_MAX_BITS));
  CUDA_TRY(cudaMalloc(&c->d_counts, sizeof(unsigned long long) * (size_t)world_size * (world_size + 1)));
  CUDA_TRY(cudaMalloc(&c->d_lut, (size_t)1 << MGPU_MAX_BITS));
  CUDA_TRY(cudaMalloc(&c->d_bin_base, sizeof(uint64_t) * RADIX));
  CUDA_TRY(cudaMalloc(&c->d_plan, sizeof(Plan)));
  *out = c;
  return 0;
}

int b200sort_mgpu_comm_destroy(b200sort_comm *c) {
  using namespace b200sort;
  if (!c) return 0;
  if (c->comm) nccl_api().CommDestroy(c->comm);
  cudaFree(c->d_hist); cudaFree(c->d_counts); cudaFree(c->d_lut); cudaFree(c->d_bin_base); cudaFree(c->d_plan);
  delete c;
  return 0;
}

int b200sort_mgpu_splitters(const uint64_t *global_hist, int bits, int world_size, uint32_t *out_bounds) {
  using namespace b200sort;
  if (!global_hist || !out_bounds || bits < 1 || bits > MGPU_MAX_BITS || world_size < 1 || world_size > RADIX)
    return fail(B200SORT_EINVAL, "bad splitter arguments");
  compute_splitters(global_hist, bits, world_size, out_bounds);
  return 0;
}

int b200sort_mgpu_plan(const uint64_t *local_hist, int bits, int world_size, const uint32_t *bounds,
                       uint64_t *out_send_counts) {
  using namespace b200sort;
  if (!local_hist || !bounds || !out_send_counts || bits < 1 || bits > MGPU_MAX_BITS || world_size < 1)
    return fail(B200SORT_EINVAL, "bad plan arguments");
  for (int r = 0; r < world_size; r++) {
    uint64_t s = 0;
    for (uint32_t b = bounds[r]; b < bounds[r + 1]; b++) s += local_hist[b];
    out_send_counts[r] = s;
  }
  return 0;
}

int b200sort_mgpu_sort_soa(b200sort_comm *c, void *keys, int key_type, int64_t num_local, int64_t capacity,
                           int ascending, int n_payloads, void *const *payloads, const uint32_t *payload_elem_bytes,
                           int64_t *out_num_local, void *stream_v) {
  using namespace b200sort;
  NcclApi &api = nccl_api();
  if (!api.ok) return fail(B200SORT_ENCCL, "NCCL unavailable: %s", api.why.c_str());
  if (!c || !out_num_local) return fail(B200SORT_EINVAL, "comm/out_num_local is NULL");
  if (capacity < num_local) return fail(B200SORT_EINVAL, "capacity smaller than num_local");
  std::vector<StreamDesc> streams;
  if (int rc = check_soa(keys, key_type, num_local, n_payloads, payloads, payload_elem_bytes, &streams)) return rc;
  if (!keys) return fail(B200SORT_EINVAL, "keys is NULL");
  cudaStream_t stream = (cudaStream_t)stream_v;
  const int kb = key_bytes_of(key_type);
  const int world = c->world;
  const int bits = std::min(MGPU_MAX_BITS, 8 * kb);
  const uint32_t nb = 1u << bits;
  const int shift = 8 * kb - bits;
  DevInfo di;
  if (int rc = dev_info(c->dev, &di)) return rc;
  const KeyOrder ko = make_key_order(key_type, ascending != 0);

  // 1-2: local histogram of the top bits, all-reduce
  CUDA_TRY(cudaMemsetAsync(c->d_hist, 0, sizeof(unsigned long long) * nb, stream));
  if (num_local > 0) {
    TopHistArgs ha{(const unsigned char *)keys, (uint32_t)kb, num_local, ko, shift, c->d_hist};
    CUDA_TRY(launch_top_hist(kb, ha, di.sm_count, stream));
  }
  std::vector<uint64_t> local_hist(nb), global_hist(nb);
  CUDA_TRY(cudaMemcpyAsync(local_hist.data(), c->d_hist, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, stream));
  NCCL_TRY(api.AllReduce(c->d_hist, c->d_hist, nb, ncclUint64, ncclSum, c->comm, stream));
  CUDA_TRY(cudaMemcpyAsync(global_hist.data(), c->d_hist, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));

  // splitters (same on every rank), destination LUT, send counts
  std::vector<uint32_t> bounds(world + 1);
  compute_splitters(global_hist.data(), bits, world, bounds.data());
  std::vector<uint8_t> lut(nb);
  std::vector<uint64_t> send(world, 0);
  for (int r = 0; r < world; r++)
    for (uint32_t b = bounds[r]; b < bounds[r + 1]; b++) { lut[b] = (uint8_t)r; send[r] += local_hist[b]; }

  // 3: all-gather the send-count rows -> full matrix m[src][dst]
  CUDA_TRY(cudaMemcpyAsync(c->d_counts, send.data(), sizeof(uint64_t) * world, cudaMemcpyHostToDevice, stream));
  NCCL_TRY(api.AllGather(c->d_counts, c->d_counts + world, world, ncclUint64, c->comm, stream));
  std::vector<uint64_t> m((size_t)world * world);
  CUDA_TRY(cudaMemcpyAsync(m.data(), c->d_counts + world, sizeof(uint64_t) * world * world, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  int64_t recv_total = 0;
  std::vector<int64_t> recv_off(world), send_off(world);
  for (int s = 0; s < world; s++) { recv_off[s] = recv_total; recv_total += (int64_t)m[(size_t)s * world + c->rank]; }
  {
    int64_t o = 0;
    for (int r = 0; r < world; r++) { send_off[r] = o; o += (int64_t)send[r]; }
  }
  // every rank can evaluate every rank's receive total, so all ranks fail together
  for (int r = 0; r < world; r++) {
    int64_t t = 0;
    for (int s = 0; s < world; s++) t += (int64_t)m[(size_t)s * world + r];
    if (t > capacity) return fail(B200SORT_ENOMEM, "rank %d would receive %lld records, capacity is %lld", r, (long long)t, (long long)capacity);
  }

  // workspace: shadow of every stream sized for max(num_local, recv_total)
  const int64_t n_ws = std::max<int64_t>(std::max(num_local, recv_total), 1);
  uint32_t stage_bytes = (uint32_t)kb;
  StreamSet ss{};
  ss.n_streams = (int)streams.size();
  for (size_t s = 0; s < streams.size(); s++) {
    const uint32_t ck = chunk_for(streams[s].ptr, streams[s].elem_bytes);
    ss.streams[s].chunk_bytes = ck;
    ss.streams[s].chunks_per_elem = streams[s].elem_bytes / ck;
    ss.streams[s].buf[0] = (unsigned char *)streams[s].ptr;
    stage_bytes = std::max(stage_bytes, ck);
  }
  const int cfg = pick_tile_cfg(kb, stage_bytes, di.smem_optin);
  const TileCfg tc = kTileCfgs[cfg];
  const int tile = tc.threads * tc.ipt;
  Layout L;
  make_layout(streams, n_ws, std::min(tile, HYB_MIN_TILE), &L);
  void *ws_v = nullptr;
  if (int rc = cached_workspace(c->dev, L.total, &ws_v)) return rc;
  unsigned char *ws = (unsigned char *)ws_v;
  for (size_t s = 0; s < streams.size(); s++) ss.streams[s].buf[1] = ws + L.shadow_off[s];

  // 4: partition by destination: one scatter pass, caller arrays -> shadow
  if (num_local > 0) {
    CUDA_TRY(cudaMemsetAsync(ws + L.ctrl_off, 0, L.ctrl_bytes, stream));
    uint64_t bin_base[RADIX] = {0};
    for (int r = 0; r < RADIX; r++) bin_base[r] = r < world ? (uint64_t)send_off[r] : (uint64_t)num_local;
    Plan plan{};
    plan.final_sel = 1; plan.n_exec = 1;
    CUDA_TRY(cudaMemcpyAsync(c->d_lut, lut.data(), nb, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_bin_base, bin_base, sizeof bin_base, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_plan, &plan, sizeof plan, cudaMemcpyHostToDevice, stream));
    SweepArgs wa{};
    wa.ss = ss; wa.n = num_local; wa.ko = ko; wa.pass = 0; wa.shift = 0;
    wa.bin_base = c->d_bin_base; wa.lookback = (uint64_t *)(ws + L.lookback_off);
    wa.tile_counter = (uint32_t *)(ws + L.tilectr_off); wa.plan = c->d_plan; wa.tag = 1; wa.stage_bytes = stage_bytes;
    wa.lut = c->d_lut; wa.lut_shift = shift;
    CUDA_TRY(launch_sweep(kb, cfg, wa, (num_local + tile - 1) / tile, di.smem_optin, di.sm_count, stream));
    // bin_base / plan / lut live in the communicator and are reused: the copies above are stream-ordered,
    // but the host arrays are stack/vector memory, so wait before they go out of scope
    CUDA_TRY(cudaStreamSynchronize(stream));
  }

  // 5: all-to-all-v, shadow -> caller arrays
  NCCL_TRY(api.GroupStart());
  for (size_t s = 0; s < streams.size(); s++) {
    const size_t eb = streams[s].elem_bytes;
    for (int p = 0; p < world; p++) {
      const size_t sb = (size_t)send[p] * eb, rb = (size_t)m[(size_t)p * world + c->rank] * eb;
      if (sb) NCCL_TRY(api.Send(ss.streams[s].buf[1] + (size_t)send_off[p] * eb, sb, ncclUint8, p, c->comm, stream));
      if (rb) NCCL_TRY(api.Recv(ss.streams[s].buf[0] + (size_t)recv_off[p] * eb, rb, ncclUint8, p, c->comm, stream));
    }
  }
  NCCL_TRY(api.GroupEnd());

  // 6: local sort of what arrived
  *out_num_local = recv_total;
  if (recv_total > 1) {
    int rc = sort_device(key_type, ascending != 0, recv_total, streams, stream, nullptr, 0);
    if (rc != 0) return rc;
  }
  return 0;
}

}  // extern "C"
