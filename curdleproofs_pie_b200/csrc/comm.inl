// The data-path collective of the hot path, INSIDE the library: NCCL over NVLink on device buffers, no torch, no
// host bounce.  Included at the end of cpg_api.cu (same translation unit).
//
// Where the path has a real exchange step:
//   * ONE large MSM (G1Point.multiexp_unchecked, stub :28, at n >= 2^16): the W Pippenger windows are split across the
//     ranks, every rank computes the Jacobian window sums of its slice, ONE ncclAllGather of W x 144 B gives every
//     rank all of them, every rank finishes with the Horner pass                       -> cpg_g1_msm_sharded
//   * ONE large proof (BASELINE config 5): the leaves (CRS bases, trackers) are split across the ranks, every rank
//     evaluates its part of each round's outputs, one all-gather of <= 10 partial sums per round  -> prove.inl / verify.inl
// Batched prove / verify shard by proof and need no collective at all.
//
// NCCL is reached through dlopen("libnccl.so.2") rather than a link-time dependency: a process that also holds
// torch (bench.py's timing barrier) has torch's bundled NCCL mapped under the same soname, and the loader then hands
// back that very copy - one NCCL per process - while a torch-free caller gets the system library.
#ifndef CPG_HOST_EMU
#include <dlfcn.h>
#include <nccl.h>
#endif

namespace {

int g_comm_rank = 0, g_comm_world = 1;

#ifndef CPG_HOST_EMU
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;
ncclComm_t g_comm = nullptr;

int nccl_load() {
    if (g_nccl.h) return 0;
    const char* names[] = {getenv("CPG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) if (nm && *nm && (h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) return fail(std::string("cpg_comm: cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "?"));
    NcclApi a; a.h = h;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString) return fail("cpg_comm: libnccl lacks a required symbol");
    g_nccl = a;
    return 0;
}
int nck(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return 0;
    return fail(std::string(what) + ": " + g_nccl.GetErrorString(r));
}
#define NCK(x) do { if (int rc_ = nck((x), #x)) return rc_; } while (0)
#endif

// all-gather of `bytes` per rank on the current stream (device buffers); world = 1: a copy
int comm_allgather(const void* d_send, void* d_recv, size_t bytes) {
    if (g_comm_world == 1) return d_send == d_recv ? 0 : cpg_d2d(d_recv, d_send, bytes);
#ifndef CPG_HOST_EMU
    NCK(g_nccl.AllGather(d_send, d_recv, bytes, ncclUint8, g_comm, cur()));
    g_launches++;
    return 0;
#else
    return fail("cpg_comm: collectives need the CUDA build");
#endif
}

// contiguous block of `total` items owned by `rank` (sizes differ by at most one) - sharding.py::shard_range
void comm_block(size_t total, int rank, int world, size_t* lo, size_t* hi) {
    size_t base = total / (size_t)world, extra = total % (size_t)world, r = (size_t)rank;
    *lo = r * base + (r < extra ? r : extra);
    *hi = *lo + base + (r < extra ? 1 : 0);
}

// How ONE proof is split over the ranks (BASELINE config 5): every rank holds the whole proof's inputs and runs the
// (cheap, deterministic) transcript and Fr vector work in full, but evaluates each round's MSMs only over ITS block of
// the leaves - CRS bases through a fixed-base table of that block only, trackers / proof points through the bucket
// method - and one all-gather per round exchanges the partial sums (<= 16 Jacobian points per proof).
// `virt`: all ranks' blocks are evaluated one after the other in THIS process - the CPU test tier's stand-in for the
// communicator (CPG_TEST_VIRTUAL_RANKS=k, honoured only while no communicator exists); the block logic and the
// summation are the product's own.
struct Shard {
    int world = 1, rank = 0; bool virt = false;
    bool on() const { return world > 1; }
    int first() const { return virt ? 0 : rank; }
    int last() const { return virt ? world : rank + 1; }
};
Shard shard_now() {
    Shard s;
    if (g_comm_world > 1) { s.world = g_comm_world; s.rank = g_comm_rank; return s; }
    if (const char* e = getenv("CPG_TEST_VIRTUAL_RANKS")) { int k = atoi(e); if (k > 1 && k <= 64) { s.world = k; s.virt = true; } }
    return s;
}
// gather = [world][cnt] partial sums, this rank's block already in place; afterwards every block is (in-place all-gather)
int shard_exchange(const Shard& sd, Jac* gather, size_t cnt) {
    if (!sd.on() || sd.virt) return 0;
    return comm_allgather(gather + (size_t)sd.rank * cnt, gather, cnt * sizeof(Jac));
}

}  // namespace

extern "C" {

int cpg_comm_unique_id(uint8_t* out128) {
#ifndef CPG_HOST_EMU
    if (int rc = nccl_load()) return rc;
    ncclUniqueId id;
    NCK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof id == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return 0;
#else
    memset(out128, 0, 128);
    return 0;
#endif
}

int cpg_comm_init(int rank, int world, const uint8_t* id128) {
    NEED_INIT();
    if (world < 1 || rank < 0 || rank >= world) return fail("cpg_comm_init: bad rank / world");
    if (g_comm_world != 1) return fail("cpg_comm_init: a communicator already exists (cpg_comm_free first)");
    if (world == 1) return 0;
#ifndef CPG_HOST_EMU
    if (!id128) return fail("cpg_comm_init: null id");
    if (int rc = nccl_load()) return rc;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NCK(g_nccl.CommInitRank(&g_comm, world, id, rank));
    g_comm_rank = rank; g_comm_world = world;
    return 0;
#else
    (void)id128;
    return fail("cpg_comm_init: multi-rank communicators need the CUDA build (the CPU test tier injects its own transport)");
#endif
}

int cpg_comm_free(void) {
#ifndef CPG_HOST_EMU
    if (g_comm) { cudaDeviceSynchronize(); g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
#endif
    g_comm_rank = 0; g_comm_world = 1;
    return 0;
}
int cpg_comm_rank(void) { return g_comm_rank; }
int cpg_comm_world(void) { return g_comm_world; }
int cpg_comm_nccl_version(void) {
#ifndef CPG_HOST_EMU
    int v = 0;
    if (nccl_load() || !g_nccl.GetVersion || g_nccl.GetVersion(&v) != ncclSuccess) return 0;
    return v;
#else
    return 0;
#endif
}

int cpg_comm_allgather(const void* d_send, void* d_recv, size_t bytes_per_rank) {
    NEED_INIT();
    return comm_allgather(d_send, d_recv, bytes_per_rank);
}

/* One n-term MSM over the communicator: windows [lo, hi) of this rank -> window sums -> ONE ncclAllGather of
 * ceil(W / world) x 144 B per rank -> Horner on every rank.  Every rank holds all bases and scalars and passes the
 * same n and window (0 = cpg_msm_pick_window(n), a function of n only).  world = 1: the plain single-GPU MSM. */
int cpg_g1_msm_sharded(const void* d_bases_aff, const uint8_t* d_scalars, size_t n, int window, void* d_out_jac) {
    NEED_INIT();
    const int world = g_comm_world, rank = g_comm_rank;
    const int c = window > 0 ? window : cpg_msm_pick_window(n);
    if (world == 1 || n == 0) return cpg_g1_msm_batched(d_bases_aff, 0, d_scalars, 1, n, c, d_out_jac);
    const size_t W = windows_for((uint32_t)c);
    const size_t width = (W + world - 1) / world;
    size_t lo, hi;
    comm_block(W, rank, world, &lo, &hi);
    Scratch sc;
    Jac* mine = sc.get<Jac>(width);
    Jac* all = sc.get<Jac>(width * world);
    Jac* wsums = sc.get<Jac>(W);
    if (!mine || !all || !wsums) return fail("cpg_g1_msm_sharded: scratch allocation failed");
    if (hi > lo) if (int rc = cpg_g1_msm_window_sums(d_bases_aff, d_scalars, n, c, (int)lo, (int)hi, mine)) return rc;
    if (int rc = comm_allgather(mine, all, width * sizeof(Jac))) return rc;
    for (int r = 0; r < world; r++) {                       // drop the padding slots: rank r's block holds its hi - lo sums
        size_t l, h;
        comm_block(W, r, world, &l, &h);
        if (h > l) if (int rc = cpg_d2d(wsums + l, all + (size_t)r * width, (h - l) * sizeof(Jac))) return rc;
    }
    return cpg_g1_msm_combine_windows(wsums, c, d_out_jac);
}

}  // extern "C"
