/*
 * comm.cu -- inter-GPU plumbing: what MPI does in the reference.
 *
 *   halo / transfer messages  MPI_Irecv/Isend/Waitall  exchange_boundary.c:33-97, restriction.c:128-192,
 *                                                      interpolation_v2.c:235-300, interpolation_v4.c:301-366
 *   norm / dot / mean / eig   MPI_Allreduce            misc.c:276,324,373, rebuild.c:195
 *
 * One process per GPU (torchrun).  Messages are grouped ncclSend/ncclRecv on the compute stream --
 * one message per neighbour rank per operation, concatenated in the reference's buffer order
 * (level.c:79-92) -- so they are stream-ordered with the pack/unpack kernels and need no host
 * wait.  The 8-byte reductions are ncclAllReduce on a device scalar slot.  NCCL is dlopen'ed when a
 * multi-rank communicator is installed, so a single-GPU process has no NCCL dependency at all.
 * The unique id is distributed through the host-provided allgather callback (torch.distributed).
 */
#include <dlfcn.h>
#include <string.h>
#include "common.cuh"

/* the slice of nccl.h we use (NCCL >= 2.7 ABI) */
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
enum { ncclSum = 0, ncclMax = 2 };
enum { ncclFloat64 = 8 };
struct NcclApi {
  void *lib;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi N = {};
static void p2p_setup(void);
extern "C" void hpgmg_b200_p2p_finalize(void);
static ncclComm_t g_comm = NULL;
static int g_rank = 0, g_nranks = 1;
static hpgmg_allgather_fn g_allgather = NULL;
static hpgmg_barrier_fn g_barrier = NULL;
static void *g_comm_ctx = NULL;

#define NCCL_CHECK(x)                                                                              \
  do {                                                                                             \
    ncclResult_t r_ = (x);                                                                         \
    if (r_ != ncclSuccess) {                                                                       \
      fprintf(stderr, "hpgmg_b200: NCCL error %d (%s) at %s:%d in `%s`\n", (int)r_,                \
              N.GetErrorString ? N.GetErrorString(r_) : "?", __FILE__, __LINE__, #x);              \
      abort();                                                                                     \
    }                                                                                              \
  } while (0)

static int load_nccl(void)
{
  if (N.lib) return 0;
  const char *names[] = { "libnccl.so.2", "libnccl.so", NULL };
  for (int i = 0; names[i] && !N.lib; i++) N.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!N.lib) { fprintf(stderr, "hpgmg_b200: cannot dlopen libnccl.so.2: %s\n", dlerror()); return 1; }
#define SYM(field, name) *(void **)(&N.field) = dlsym(N.lib, name); if (!N.field) { fprintf(stderr, "hpgmg_b200: %s missing in libnccl\n", name); return 2; }
  SYM(GetUniqueId, "ncclGetUniqueId");  SYM(CommInitRank, "ncclCommInitRank");  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");      SYM(Send, "ncclSend");                  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");    SYM(GroupEnd, "ncclGroupEnd");          SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return 0;
}

extern "C" void hpgmg_b200_set_comm(int my_rank, int num_ranks, hpgmg_allgather_fn allgather, hpgmg_barrier_fn barrier, void *ctx)
{
  g_rank = my_rank;  g_nranks = num_ranks;  g_allgather = allgather;  g_barrier = barrier;  g_comm_ctx = ctx;
  if (num_ranks <= 1) return;
  if (!allgather) { fprintf(stderr, "hpgmg_b200_set_comm: an allgather callback is required for num_ranks>1\n"); exit(1); }
  hpgmg_rt_scalar_slots();                                   /* forces device init */
  if (load_nccl() != 0) exit(1);
  ncclUniqueId id;
  memset(&id, 0, sizeof(id));
  if (my_rank == 0) NCCL_CHECK(N.GetUniqueId(&id));
  ncclUniqueId *all = (ncclUniqueId *)malloc(sizeof(ncclUniqueId) * (size_t)num_ranks);
  allgather(&id, all, sizeof(ncclUniqueId), ctx);
  id = all[0];
  free(all);
  NCCL_CHECK(N.CommInitRank(&g_comm, num_ranks, id, my_rank));
  p2p_setup();
}

extern "C" void hpgmg_b200_comm_finalize(void)
{
  hpgmg_b200_p2p_finalize();
  if (g_comm) { hpgmg_rt_sync(); N.CommDestroy(g_comm); g_comm = NULL; }
  g_nranks = 1;  g_rank = 0;
}

extern "C" int hpgmg_comm_rank(void) { return g_rank; }
extern "C" int hpgmg_comm_size(void) { return g_nranks; }
extern "C" void hpgmg_comm_barrier(void) { if (g_nranks > 1 && g_barrier) { hpgmg_rt_sync(); g_barrier(g_comm_ctx); } }

/* ---- reductions --------------------------------------------------------------------------------- */
/* MPI_Allreduce of ONE double over the ranks that share a level (misc.c:276,324,373 on level->MPI_COMM_ALLREDUCE, the
 * sub-communicator of mg.c:979-997: ranks 0 .. num_ranks-1 of the level) or over all ranks (rebuild.c:195).
 *
 * With peer memory: every participant stores its value, tagged with the reduction's sequence number, into a mailbox in
 * each participant's arena (the 16-byte LL slot of p2p.cuh: value and flag arrive together) and then polls its own mailbox
 * until all values are there; they are combined in rank order, so every rank gets the same bits.  One kernel of one warp,
 * sequence numbers in device memory: capturable, no host involvement, a few microseconds instead of an ncclAllReduce launch.
 * Each subset size has its own mailboxes and counter, so ranks outside a sub-communicator never fall out of step.  Two
 * slot sets alternate by sequence parity: a rank can only be one reduction ahead of a peer, because finishing reduction k
 * needs the peer's value k, which the peer sends after it has finished k-1. */
#include "p2p.cuh"
#define AR_MAX_RANKS 32
struct ARPeers { uint4 *mbox[AR_MAX_RANKS]; };          /* rank p's mailbox array, as mapped into this process */
static uint4 *g_ar_mbox = NULL;                           /* mine: [subset size][parity][sender] */
static unsigned long long *g_ar_seq = NULL;               /* device: reductions completed, per subset size */
static ARPeers g_ar_peers;
__host__ __device__ static inline size_t ar_slot(const int n, const int parity, const int sender) { return ((size_t)(n * 2 + parity)) * AR_MAX_RANKS + sender; }

__global__ void ll_allreduce_kernel(double *value, const ARPeers P, const int n, const int my_rank, const int op, unsigned long long *seq)
{
  PDL_WAIT();
  __shared__ double vals[AR_MAX_RANKS];
  const unsigned long long k = *(volatile unsigned long long *)seq;
  const unsigned flag = (unsigned)(k + 1);
  const int parity = (int)(k & 1), t = threadIdx.x;
  const double mine = *(volatile double *)value;
  if (t < n) ll_store(P.mbox[t] + ar_slot(n, parity, my_rank), mine, flag);
  if (t < n) vals[t] = ll_load(P.mbox[my_rank] + ar_slot(n, parity, t), flag);
  __syncwarp();
  if (t == 0) {
    double r = vals[0];
    for (int q = 1; q < n; q++) r = op ? r + vals[q] : (vals[q] > r ? vals[q] : r);
    *value = r;
    *seq = k + 1;
  }
}

static int g_p2p_enabled = 0;
static void require_comm(const char *what)
{
  if (!g_comm) { fprintf(stderr, "hpgmg_b200: %s across ranks but no communicator was installed (hpgmg_b200_set_comm)\n", what); exit(1); }
}
/* reduce the device scalar `s` in place over ranks 0..n-1 (op 0: max, 1: sum); stream-ordered */
static void allreduce_device_scalar(double *s, const int n, const int op, const char *what)
{
  if (n <= 1 || g_nranks <= 1 || g_rank >= n) return;       /* ranks outside the sub-communicator keep their local value */
  require_comm(what);
  if (g_p2p_enabled && n <= AR_MAX_RANKS) {
    LAUNCH(ll_allreduce_kernel, 1, 32, 0, s, g_ar_peers, n, g_rank, op, g_ar_seq + n);
    return;
  }
  if (n != g_nranks) {
    fprintf(stderr, "hpgmg_b200: %s over %d of %d ranks needs the peer-memory path (CUDA IPC unavailable here)\n", what, n, g_nranks);
    exit(1);
  }
  NCCL_CHECK(N.AllReduce(s, s, 1, ncclFloat64, op ? ncclSum : ncclMax, g_comm, g_stream));
  g_launches++;
}

extern "C" void hpgmg_comm_allreduce_slot_max(level_type *level, int slot)
{
  allreduce_device_scalar(hpgmg_rt_scalar_slots() + slot, level->num_ranks, 0, "norm");
}

static double allreduce_host_value(level_type *level, double v, int op, const char *what, int world)
{
  const int n = world ? g_nranks : level->num_ranks;
  if (g_nranks <= 1 || n <= 1 || g_rank >= n) return v;
  const int slot = HPGMG_SLOT_SCRATCH + 3;
  double *s = hpgmg_rt_scalar_slots() + slot;
  CUDA_CHECK(cudaMemcpyAsync(s, &v, sizeof(double), cudaMemcpyHostToDevice, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  allreduce_device_scalar(s, n, op, what);
  double r = 0.0;
  hpgmg_rt_read_scalars(&r, slot, 1);
  return r;
}
/* over the ranks that share the level (MPI_COMM_ALLREDUCE of the reference) */
extern "C" double hpgmg_comm_allreduce_max(level_type *level, double v) { return allreduce_host_value(level, v, 0, "max-reduction", 0); }
extern "C" double hpgmg_comm_allreduce_sum(level_type *level, double v) { return allreduce_host_value(level, v, 1, "sum-reduction", 0); }
/* over ALL ranks, whatever the level's rank count: the reference reduces lambda_max on MPI_COMM_WORLD (rebuild.c:195) */
extern "C" double hpgmg_comm_allreduce_max_world(level_type *level, double v) { return allreduce_host_value(level, v, 0, "max-reduction (world)", 1); }

/* ---- point-to-point ------------------------------------------------------------------------------ */
/* ghost exchange of one level: receive into recv_buffers, send from send_buffers (both sides use the
 * same communicator_type) */
extern "C" void hpgmg_comm_exchange(level_type *level, communicator_type *C, int tag)
{
  (void)tag;
  if (C->num_recvs + C->num_sends == 0) return;
  (void)level;                              /* ranks are world ranks on every level (rank_of_box) */
  if (!g_comm) { fprintf(stderr, "hpgmg_b200: ghost exchange with other ranks needs a communicator (hpgmg_b200_set_comm)\n"); exit(1); }
  NCCL_CHECK(N.GroupStart());
  for (int n = 0; n < C->num_recvs; n++) NCCL_CHECK(N.Recv(C->recv_buffers[n], (size_t)C->recv_sizes[n], ncclFloat64, C->recv_ranks[n], g_comm, g_stream));
  for (int n = 0; n < C->num_sends; n++) NCCL_CHECK(N.Send(C->send_buffers[n], (size_t)C->send_sizes[n], ncclFloat64, C->send_ranks[n], g_comm, g_stream));
  NCCL_CHECK(N.GroupEnd());
  g_launches++;
}
extern "C" void hpgmg_comm_exchange_wait(level_type *level, communicator_type *C) { (void)level; (void)C; /* stream-ordered */ }

/* inter-level transfer: the sender's list lives on one level, the receiver's on the other */
extern "C" void hpgmg_comm_transfer(level_type *level_send, communicator_type *Cs, level_type *level_recv, communicator_type *Cr, int tag)
{
  (void)tag; (void)level_send; (void)level_recv;
  if (Cr->num_recvs + Cs->num_sends == 0) return;
  if (!g_comm) { fprintf(stderr, "hpgmg_b200: inter-level transfer needs a communicator\n"); exit(1); }
  NCCL_CHECK(N.GroupStart());
  for (int n = 0; n < Cr->num_recvs; n++) NCCL_CHECK(N.Recv(Cr->recv_buffers[n], (size_t)Cr->recv_sizes[n], ncclFloat64, Cr->recv_ranks[n], g_comm, g_stream));
  for (int n = 0; n < Cs->num_sends; n++) NCCL_CHECK(N.Send(Cs->send_buffers[n], (size_t)Cs->send_sizes[n], ncclFloat64, Cs->send_ranks[n], g_comm, g_stream));
  NCCL_CHECK(N.GroupEnd());
  g_launches++;
}
extern "C" void hpgmg_comm_transfer_wait(level_type *level_send, communicator_type *Cs, level_type *level_recv, communicator_type *Cr)
{ (void)level_send; (void)Cs; (void)level_recv; (void)Cr; /* stream-ordered */ }

/* ================================================================================================
 * Direct peer-to-peer ghost exchange over NVLink (replaces the ncclSend/ncclRecv pair, whose ~20 us
 * of launch latency per message group dominates a multigrid cycle made of ~400 tiny exchanges).
 *
 * Every rank owns one "comm arena" of device memory, exported once with CUDA IPC and mapped by all
 * peers.  The slot arrays of the LL protocol (p2p.cuh) live in it: the SENDER's pack blocks store
 * their faces/edges/corners straight into the RECEIVER's slots, the receiver's unpack blocks poll
 * them -- inside the same kernel that does the GPU-local part of the ghost fill (ghost.cu).  Sequence
 * numbers live in device memory and are advanced by the kernel itself, so everything is capturable in
 * the solve's CUDA graph; the host never waits.  Message contents and order are the reference's
 * (level.c:79-92, 724, 878).
 * ================================================================================================ */
#include <map>
#include <vector>

struct P2PHost {
  P2PPlan *plan;                                   /* device */
  blockCopy_type *pack, *unpack;                   /* device copies: write.ptr -> remote buffers; subtype = neighbour index */
  int npack, nunpack;
};

static char *g_arena = NULL;
static size_t g_arena_size = 0, g_arena_used = 0;
static std::vector<char *> g_peer_arena;          /* my mapping of every rank's arena */
static std::map<communicator_type *, P2PHost> g_p2p;
struct XferPlan;
struct XferHost { XferPlan *plan; blockCopy_type *pack; int npack; };          /* inter-level transfers over peer memory (below) */
static std::map<communicator_type *, XferHost> g_xfer_send, g_xfer_recv;
static size_t g_arena_base = 0;                   /* permanent allocations (reduction mailboxes) end here */
static long g_arena_live = 0;                     /* allocations handed out and not yet freed */

extern "C" int hpgmg_rt_is_comm_memory(const void *p) { return g_arena && (const char *)p >= g_arena && (const char *)p < g_arena + g_arena_size; }

extern "C" void *hpgmg_rt_alloc_comm(size_t bytes)
{
  if (!g_arena) return hpgmg_rt_alloc_zero(bytes);
  const size_t need = (bytes + 255) & ~(size_t)255;
  if (g_arena_used + need > g_arena_size) {
    fprintf(stderr, "hpgmg_b200: comm arena exhausted (%zu of %zu bytes used, %zu requested); set HPGMG_B200_COMM_ARENA_MB\n", g_arena_used, g_arena_size, bytes);
    exit(1);
  }
  void *p = g_arena + g_arena_used;               /* bump allocation out of memory that was zeroed synchronously (at creation or at the last
                                                     collective reset): a stream-ordered memset here could run AFTER a fast peer's first store */
  g_arena_used += need;
  g_arena_live++;
  return p;
}
extern "C" void hpgmg_rt_free_comm(void *p) { (void)p; if (g_arena_live > 0) g_arena_live--; }

/* Collective (called by every rank at the start of create_level): once NO rank holds live arena memory any more -- all
 * hierarchies destroyed -- the arena is zeroed and handed out from the start again, so a process that builds and destroys
 * many hierarchies does not run out.  Zeroing matters: LL slots are recognised by their sequence flags. */
extern "C" void hpgmg_comm_recycle_arena(void)
{
  if (!g_arena || g_nranks <= 1 || !g_allgather) return;
  int mine = (g_arena_live == 0 && g_p2p.empty() && g_xfer_send.empty() && g_xfer_recv.empty()) ? 1 : 0, all_free = 1;
  std::vector<int> everyone((size_t)g_nranks);
  g_allgather(&mine, everyone.data(), sizeof(int), g_comm_ctx);
  for (int r = 0; r < g_nranks; r++) all_free &= everyone[r];
  if (!all_free || g_arena_used == g_arena_base) return;
  hpgmg_rt_sync();
  if (g_barrier) g_barrier(g_comm_ctx);             /* nobody is still storing into a peer's old slots */
  CUDA_CHECK(cudaMemset(g_arena + g_arena_base, 0, g_arena_used - g_arena_base));
  g_arena_used = g_arena_base;
  if (g_barrier) g_barrier(g_comm_ctx);
}

/* called from hpgmg_b200_set_comm once NCCL is up: create, export and map the arenas */
static void p2p_setup(void)
{
  const char *off = getenv("HPGMG_B200_NO_P2P");
  if (off && atoi(off)) return;
  const char *mb = getenv("HPGMG_B200_COMM_ARENA_MB");
  g_arena_size = (size_t)(mb ? atol(mb) : 512) << 20;
  if (cudaMalloc((void **)&g_arena, g_arena_size) != cudaSuccess) { cudaGetLastError(); g_arena = NULL; return; }
  CUDA_CHECK(cudaMemset(g_arena, 0, g_arena_size));
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, g_arena) != cudaSuccess) { cudaGetLastError(); fprintf(stderr, "hpgmg_b200: CUDA IPC export failed; falling back to NCCL send/recv\n"); cudaFree(g_arena); g_arena = NULL; return; }
  std::vector<cudaIpcMemHandle_t> all((size_t)g_nranks);
  g_allgather(&mine, all.data(), sizeof(mine), g_comm_ctx);
  g_peer_arena.assign((size_t)g_nranks, (char *)NULL);
  int ok = 1;
  for (int r = 0; r < g_nranks; r++) {
    if (r == g_rank) { g_peer_arena[r] = g_arena; continue; }
    void *p = NULL;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    g_peer_arena[r] = (char *)p;
  }
  /* everybody must agree, otherwise the protocol deadlocks */
  std::vector<int> oks((size_t)g_nranks);
  g_allgather(&ok, oks.data(), sizeof(int), g_comm_ctx);
  for (int r = 0; r < g_nranks; r++) ok &= oks[r];
  if (!ok) { if (g_rank == 0) fprintf(stderr, "hpgmg_b200: CUDA IPC peer mapping unavailable; using NCCL send/recv for halos\n"); g_p2p_enabled = 0; return; }
  /* reduction mailboxes: the first thing in every rank's arena, hence at the same offset everywhere */
  const size_t mbox_bytes = sizeof(uint4) * ar_slot(AR_MAX_RANKS + 1, 0, 0);
  g_ar_mbox = (uint4 *)hpgmg_rt_alloc_comm(mbox_bytes);
  g_arena_live--;                                    /* permanent */
  g_arena_base = g_arena_used;
  for (int r = 0; r < AR_MAX_RANKS; r++) g_ar_peers.mbox[r] = (r < g_nranks) ? (uint4 *)(g_peer_arena[r] + ((char *)g_ar_mbox - g_arena)) : NULL;
  CUDA_CHECK(cudaMalloc(&g_ar_seq, sizeof(unsigned long long) * (AR_MAX_RANKS + 1)));
  CUDA_CHECK(cudaMemset(g_ar_seq, 0, sizeof(unsigned long long) * (AR_MAX_RANKS + 1)));
  if (g_barrier) g_barrier(g_comm_ctx);             /* every arena is zeroed and mapped before anybody stores into one */
  g_p2p_enabled = 1;
}

struct P2PWire {                                   /* what a rank tells the others about one communicator */
  int nrecv;
  int recv_from[P2P_MAX_NEIGHBOURS], recv_size[P2P_MAX_NEIGHBOURS];
  long long ll_off[P2P_MAX_NEIGHBOURS];            /* offset of the 2 x recv_size slot arrays in my arena */
};

#define P2P_CHUNK_CELLS 2048
static void split_large_entries(std::vector<blockCopy_type> &list)
{
  static int chunk = -1;
  if (chunk < 0) { const char *e = getenv("HPGMG_B200_P2P_CHUNK"); chunk = e ? atoi(e) : P2P_CHUNK_CELLS; }
  if (chunk <= 0) return;
  std::vector<blockCopy_type> out;
  for (size_t e = 0; e < list.size(); e++) {
    const blockCopy_type B = list[e];
    const long cells = (long)B.dim.i * B.dim.j * B.dim.k;
    if (cells <= chunk) { out.push_back(B); continue; }
    int cj = chunk / B.dim.i;                         /* rows of a slab */
    if (cj < 1) cj = 1;
    if (cj > B.dim.j) cj = B.dim.j;
    int ck = (cj == B.dim.j) ? chunk / (B.dim.i * B.dim.j) : 1;   /* whole planes only if a plane fits */
    if (ck < 1) ck = 1;
    for (int k0 = 0; k0 < B.dim.k; k0 += ck)
      for (int j0 = 0; j0 < B.dim.j; j0 += cj) {
        blockCopy_type S = B;
        S.dim.j = (j0 + cj <= B.dim.j) ? cj : B.dim.j - j0;
        S.dim.k = (k0 + ck <= B.dim.k) ? ck : B.dim.k - k0;
        S.read.j += j0;   S.read.k += k0;
        S.write.j += j0;  S.write.k += k0;
        out.push_back(S);
      }
  }
  list.swap(out);
}

extern "C" void hpgmg_comm_register_exchange(level_type *level, int shape)
{
  if (!g_p2p_enabled || g_nranks <= 1 || hpgmg_rt_layout_only()) return;
  communicator_type *C = &level->exchange_ghosts[shape];
  P2PWire mine;
  memset(&mine, 0, sizeof(mine));
  const int fits = (C->num_recvs <= P2P_MAX_NEIGHBOURS && C->num_sends <= P2P_MAX_NEIGHBOURS);
  P2PPlan h;
  memset(&h, 0, sizeof(h));
  if (fits) {
    mine.nrecv = C->num_recvs;
    for (int n = 0; n < C->num_recvs; n++) {
      uint4 *slots = (uint4 *)hpgmg_rt_alloc_comm(sizeof(uint4) * 2 * (size_t)C->recv_sizes[n]);
      h.ll_local[n] = slots;
      h.recv_size[n] = C->recv_sizes[n];
      mine.recv_from[n] = C->recv_ranks[n];
      mine.recv_size[n] = C->recv_sizes[n];
      mine.ll_off[n] = (long long)((char *)slots - g_arena);
    }
  } else mine.nrecv = -1;
  std::vector<P2PWire> all((size_t)g_nranks);
  g_allgather(&mine, all.data(), sizeof(P2PWire), g_comm_ctx);
  for (int r = 0; r < g_nranks; r++) if (all[r].nrecv < 0) return;            /* somebody cannot: everybody keeps NCCL for this one */
  if (C->num_recvs + C->num_sends == 0) return;

  for (int n = 0; n < C->num_sends; n++) {
    const int R = C->send_ranks[n];
    int found = -1;
    for (int m = 0; m < all[R].nrecv; m++) if (all[R].recv_from[m] == g_rank) found = m;
    if (found < 0 || all[R].recv_size[found] != C->send_sizes[n]) {
      fprintf(stderr, "hpgmg_b200: rank %d expects %d doubles from rank %d, which sends %d\n", R, found < 0 ? -1 : all[R].recv_size[found], g_rank, C->send_sizes[n]);
      exit(1);
    }
    h.ll_remote[n] = (uint4 *)(g_peer_arena[R] + all[R].ll_off[found]);
    h.send_size[n] = C->send_sizes[n];
  }
  /* device copies of the pack / unpack lists with subtype = neighbour index */
  P2PHost H;
  memset(&H, 0, sizeof(H));
  std::vector<blockCopy_type> pack(C->blocks[0], C->blocks[0] + C->num_blocks[0]), unpack(C->blocks[2], C->blocks[2] + C->num_blocks[2]);
  for (size_t e = 0; e < pack.size(); e++) {
    int n = -1;
    for (int m = 0; m < C->num_sends; m++) if (pack[e].write.ptr == C->send_buffers[m]) n = m;
    if (n < 0) { fprintf(stderr, "hpgmg_b200: pack entry without a send buffer\n"); exit(1); }
    pack[e].subtype = n;
  }
  for (size_t e = 0; e < unpack.size(); e++) {
    int n = -1;
    for (int m = 0; m < C->num_recvs; m++) if (unpack[e].read.ptr == C->recv_buffers[m]) n = m;
    if (n < 0) { fprintf(stderr, "hpgmg_b200: unpack entry without a receive buffer\n"); exit(1); }
    unpack[e].subtype = n;
  }
  /* One thread block works off one list entry.  A face of a 128^3 box is 32768 cells: 24 such blocks per fill would carry
   * the whole NVLink traffic of a rank on 24 of its 148 SMs.  Cut large entries into slabs of <= P2P_CHUNK_CELLS along j
   * (then k); an entry's read and write side index the same (i,j,k), so a slab is the entry with shifted origins. */
  split_large_entries(pack);
  split_large_entries(unpack);
  H.npack = (int)pack.size();  H.nunpack = (int)unpack.size();
  if (H.npack) { CUDA_CHECK(cudaMalloc(&H.pack, pack.size() * sizeof(blockCopy_type))); CUDA_CHECK(cudaMemcpy(H.pack, pack.data(), pack.size() * sizeof(blockCopy_type), cudaMemcpyHostToDevice)); }
  if (H.nunpack) { CUDA_CHECK(cudaMalloc(&H.unpack, unpack.size() * sizeof(blockCopy_type))); CUDA_CHECK(cudaMemcpy(H.unpack, unpack.data(), unpack.size() * sizeof(blockCopy_type), cudaMemcpyHostToDevice)); }
  CUDA_CHECK(cudaMalloc(&H.plan, sizeof(P2PPlan)));
  CUDA_CHECK(cudaMemcpy(H.plan, &h, sizeof(P2PPlan), cudaMemcpyHostToDevice));
  g_p2p[C] = H;
}

extern "C" void hpgmg_comm_unregister(communicator_type *C)
{
  std::map<communicator_type *, P2PHost>::iterator it = g_p2p.find(C);
  if (it == g_p2p.end()) return;
  hpgmg_rt_sync();
  if (it->second.pack) cudaFree(it->second.pack);
  if (it->second.unpack) cudaFree(it->second.unpack);
  cudaFree(it->second.plan);
  g_p2p.erase(it);
}

/* hand the peer-exchange state of a communicator to the kernels in ghost.cu; 0 if it is not on the peer path */
int hpgmg_comm_p2p_lookup(level_type *level, int shape, const blockCopy_type **pack, int *npack, const blockCopy_type **unpack, int *nunpack, P2PPlan **plan)
{
  if (!g_p2p_enabled) return 0;
  std::map<communicator_type *, P2PHost>::iterator it = g_p2p.find(&level->exchange_ghosts[shape]);
  if (it == g_p2p.end()) return 0;
  *pack = it->second.pack;      *npack = it->second.npack;
  *unpack = it->second.unpack;  *nunpack = it->second.nunpack;
  *plan = it->second.plan;
  return 1;
}

/* ================================================================================================
 * Inter-level transfers (restriction / interpolation where box ownership changes) over peer memory.
 *
 * One message per (sender, receiver) pair per transfer, as in the reference (restriction.c:128-192,
 * interpolation_v2.c:235-300): the sender's pack kernel writes the message straight into the RECEIVER's receive
 * buffer (the pack list is re-pointed at the peer mapping of that buffer), then a one-warp kernel publishes it:
 * system-scope fence, then the message number into a flag word next to the buffer.  The receiver's wait kernel
 * spins on the flag, its unpack kernel reads local memory, and an ack kernel stores the message number back into
 * the sender's arena; the sender does not overwrite the (single) buffer before it has seen the ack of the previous
 * message -- which in a V- or F-cycle arrived long ago, because restriction and interpolation between two levels
 * alternate.  All counters live in device memory and are advanced by the kernels, so the whole exchange is
 * recorded into the solve's CUDA graph; versus grouped ncclSend/ncclRecv it removes ~20 us per transfer.
 * ================================================================================================ */
struct XferPlan {                                  /* device-resident; one per communicator SIDE */
  unsigned long long seq;                          /* messages completed on this side */
  int n;                                           /* neighbours */
  unsigned long long *remote[P2P_MAX_NEIGHBOURS];  /* sender: receiver's flag word; receiver: sender's ack word */
  unsigned long long *local[P2P_MAX_NEIGHBOURS];   /* sender: my ack word (the receiver stores into it); receiver: my flag word */
};

__global__ void xfer_pre_kernel(XferPlan *P)       /* sender: the receivers have consumed my previous message */
{
  PDL_WAIT();
  const unsigned long long k = *(volatile unsigned long long *)&P->seq;
  if ((int)threadIdx.x < P->n) while (*(volatile unsigned long long *)P->local[threadIdx.x] < k) { }
}
__global__ void xfer_post_kernel(XferPlan *P)      /* sender: publish message k (the pack kernel before me wrote it) */
{
  PDL_WAIT();
  const unsigned long long k = *(volatile unsigned long long *)&P->seq;
  __threadfence_system();
  if ((int)threadIdx.x < P->n) *(volatile unsigned long long *)P->remote[threadIdx.x] = k + 1;
  __syncwarp();
  if (threadIdx.x == 0) P->seq = k + 1;
}
__global__ void xfer_wait_kernel(XferPlan *P)      /* receiver: message k has arrived from everybody */
{
  PDL_WAIT();
  const unsigned long long k = *(volatile unsigned long long *)&P->seq;
  if ((int)threadIdx.x < P->n) while (*(volatile unsigned long long *)P->local[threadIdx.x] < k + 1) { }
  __threadfence_system();
}
__global__ void xfer_ack_kernel(XferPlan *P)       /* receiver: the unpack kernel before me is done with the buffers */
{
  PDL_WAIT();
  const unsigned long long k = *(volatile unsigned long long *)&P->seq;
  if ((int)threadIdx.x < P->n) *(volatile unsigned long long *)P->remote[threadIdx.x] = k + 1;
  __syncwarp();
  if (threadIdx.x == 0) P->seq = k + 1;
}

struct XferWire {                                  /* what a rank tells the others about one transfer */
  int nrecv, nsend;
  int recv_from[P2P_MAX_NEIGHBOURS], recv_size[P2P_MAX_NEIGHBOURS];
  long long data_off[P2P_MAX_NEIGHBOURS], flag_off[P2P_MAX_NEIGHBOURS];     /* in the receiver's arena */
  int send_to[P2P_MAX_NEIGHBOURS];
  long long ack_off[P2P_MAX_NEIGHBOURS];                                    /* in the sender's arena */
};

/* collective: every rank calls it for every (sending communicator, receiving communicator) pair of the hierarchy, in the same order */
extern "C" void hpgmg_comm_register_transfer(communicator_type *Cs, communicator_type *Cr)
{
  if (!g_p2p_enabled || g_nranks <= 1 || hpgmg_rt_layout_only()) return;
  static int off = -1;
  if (off < 0) { const char *e = getenv("HPGMG_B200_NCCL_TRANSFERS"); off = (e && atoi(e)) ? 1 : 0; }
  if (off) return;
  const int nsend = Cs ? Cs->num_sends : 0, nrecv = Cr ? Cr->num_recvs : 0;
  XferWire mine;
  memset(&mine, 0, sizeof(mine));
  int fits = nsend <= P2P_MAX_NEIGHBOURS && nrecv <= P2P_MAX_NEIGHBOURS;
  for (int n = 0; fits && n < nrecv; n++) if (!hpgmg_rt_is_comm_memory(Cr->recv_buffers[n])) fits = 0;
  XferPlan hs, hr;
  memset(&hs, 0, sizeof(hs));  memset(&hr, 0, sizeof(hr));
  if (fits) {
    mine.nrecv = nrecv;  mine.nsend = nsend;
    for (int n = 0; n < nrecv; n++) {
      unsigned long long *flag = (unsigned long long *)hpgmg_rt_alloc_comm(sizeof(unsigned long long));
      hr.local[n] = flag;
      mine.recv_from[n] = Cr->recv_ranks[n];  mine.recv_size[n] = Cr->recv_sizes[n];
      mine.data_off[n] = (long long)((char *)Cr->recv_buffers[n] - g_arena);
      mine.flag_off[n] = (long long)((char *)flag - g_arena);
    }
    for (int n = 0; n < nsend; n++) {
      unsigned long long *ack = (unsigned long long *)hpgmg_rt_alloc_comm(sizeof(unsigned long long));
      hs.local[n] = ack;
      mine.send_to[n] = Cs->send_ranks[n];
      mine.ack_off[n] = (long long)((char *)ack - g_arena);
    }
  } else mine.nrecv = -1;
  std::vector<XferWire> all((size_t)g_nranks);
  g_allgather(&mine, all.data(), sizeof(XferWire), g_comm_ctx);
  for (int r = 0; r < g_nranks; r++) if (all[r].nrecv < 0) return;            /* somebody cannot: everybody keeps NCCL for this one */

  if (nsend > 0) {
    std::vector<blockCopy_type> pack(Cs->blocks[0], Cs->blocks[0] + Cs->num_blocks[0]);
    std::vector<char *> remote_data((size_t)nsend);
    for (int n = 0; n < nsend; n++) {
      const int R = Cs->send_ranks[n];
      int found = -1;
      for (int m = 0; m < all[R].nrecv; m++) if (all[R].recv_from[m] == g_rank) found = m;
      if (found < 0 || all[R].recv_size[found] != Cs->send_sizes[n]) {
        fprintf(stderr, "hpgmg_b200: transfer: rank %d expects %d doubles from rank %d, which sends %d\n", R, found < 0 ? -1 : all[R].recv_size[found], g_rank, Cs->send_sizes[n]);
        exit(1);
      }
      remote_data[n] = g_peer_arena[R] + all[R].data_off[found];
      hs.remote[n] = (unsigned long long *)(g_peer_arena[R] + all[R].flag_off[found]);
    }
    for (size_t e = 0; e < pack.size(); e++) {                                /* re-point the pack list at the receivers' buffers */
      int n = -1;
      for (int m = 0; m < nsend; m++)
        if ((char *)pack[e].write.ptr >= (char *)Cs->send_buffers[m] && (char *)pack[e].write.ptr < (char *)(Cs->send_buffers[m] + Cs->send_sizes[m])) n = m;
      if (n < 0) { fprintf(stderr, "hpgmg_b200: transfer pack entry without a send buffer\n"); exit(1); }
      pack[e].write.ptr = (double *)(remote_data[n] + ((char *)pack[e].write.ptr - (char *)Cs->send_buffers[n]));
    }
    XferHost H;
    memset(&H, 0, sizeof(H));
    hs.n = nsend;
    H.npack = (int)pack.size();
    if (H.npack) { CUDA_CHECK(cudaMalloc(&H.pack, pack.size() * sizeof(blockCopy_type))); CUDA_CHECK(cudaMemcpy(H.pack, pack.data(), pack.size() * sizeof(blockCopy_type), cudaMemcpyHostToDevice)); }
    CUDA_CHECK(cudaMalloc(&H.plan, sizeof(XferPlan)));
    CUDA_CHECK(cudaMemcpy(H.plan, &hs, sizeof(XferPlan), cudaMemcpyHostToDevice));
    g_xfer_send[Cs] = H;
  }
  if (nrecv > 0) {
    for (int n = 0; n < nrecv; n++) {
      const int S = Cr->recv_ranks[n];
      int found = -1;
      for (int q = 0; q < all[S].nsend; q++) if (all[S].send_to[q] == g_rank) found = q;
      if (found < 0) { fprintf(stderr, "hpgmg_b200: transfer: rank %d does not send to rank %d, which expects a message\n", S, g_rank); exit(1); }
      hr.remote[n] = (unsigned long long *)(g_peer_arena[S] + all[S].ack_off[found]);
    }
    XferHost H;
    memset(&H, 0, sizeof(H));
    hr.n = nrecv;
    CUDA_CHECK(cudaMalloc(&H.plan, sizeof(XferPlan)));
    CUDA_CHECK(cudaMemcpy(H.plan, &hr, sizeof(XferPlan), cudaMemcpyHostToDevice));
    g_xfer_recv[Cr] = H;
  }
}

extern "C" void hpgmg_comm_unregister_transfer(communicator_type *C)
{
  for (int side = 0; side < 2; side++) {
    std::map<communicator_type *, XferHost> &M = side ? g_xfer_recv : g_xfer_send;
    std::map<communicator_type *, XferHost>::iterator it = M.find(C);
    if (it == M.end()) continue;
    hpgmg_rt_sync();
    if (it->second.pack) cudaFree(it->second.pack);
    cudaFree(it->second.plan);
    M.erase(it);
  }
}

/* sender side: 1 if this communicator's messages travel over peer memory; *pack = the re-pointed pack list */
extern "C" int hpgmg_comm_xfer_send_lookup(communicator_type *Cs, const blockCopy_type **pack, int *npack)
{
  std::map<communicator_type *, XferHost>::iterator it = g_xfer_send.find(Cs);
  if (it == g_xfer_send.end()) return 0;
  *pack = it->second.pack;  *npack = it->second.npack;
  return 1;
}
extern "C" int hpgmg_comm_xfer_recv_lookup(communicator_type *Cr) { return g_xfer_recv.find(Cr) != g_xfer_recv.end(); }
extern "C" void hpgmg_comm_xfer_pre(communicator_type *Cs)  { LAUNCH(xfer_pre_kernel, 1, 32, 0, g_xfer_send[Cs].plan); }
extern "C" void hpgmg_comm_xfer_post(communicator_type *Cs) { LAUNCH(xfer_post_kernel, 1, 32, 0, g_xfer_send[Cs].plan); }
extern "C" void hpgmg_comm_xfer_wait(communicator_type *Cr) { LAUNCH(xfer_wait_kernel, 1, 32, 0, g_xfer_recv[Cr].plan); }
extern "C" void hpgmg_comm_xfer_ack(communicator_type *Cr)  { LAUNCH(xfer_ack_kernel, 1, 32, 0, g_xfer_recv[Cr].plan); }

extern "C" void hpgmg_b200_p2p_finalize(void)
{
  if (!g_arena) return;
  hpgmg_rt_sync();
  for (int r = 0; r < (int)g_peer_arena.size(); r++) if (r != g_rank && g_peer_arena[r]) cudaIpcCloseMemHandle(g_peer_arena[r]);
  g_peer_arena.clear();
  if (g_barrier) g_barrier(g_comm_ctx);             /* nobody frees while a peer may still have it mapped */
  cudaFree(g_arena);
  if (g_ar_seq) cudaFree(g_ar_seq);
  g_ar_seq = NULL;  g_ar_mbox = NULL;
  g_arena = NULL;  g_arena_size = g_arena_used = g_arena_base = 0;  g_arena_live = 0;  g_p2p_enabled = 0;
}
extern "C" int hpgmg_b200_p2p_enabled(void) { return g_p2p_enabled; }
