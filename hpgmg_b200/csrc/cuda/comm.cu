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
}

extern "C" void hpgmg_b200_comm_finalize(void)
{
  if (g_comm) { hpgmg_rt_sync(); N.CommDestroy(g_comm); g_comm = NULL; }
  g_nranks = 1;  g_rank = 0;
}

extern "C" int hpgmg_comm_rank(void) { return g_rank; }
extern "C" int hpgmg_comm_size(void) { return g_nranks; }
extern "C" void hpgmg_comm_barrier(void) { if (g_nranks > 1 && g_barrier) { hpgmg_rt_sync(); g_barrier(g_comm_ctx); } }

/* ---- reductions --------------------------------------------------------------------------------- */
static void require_world(const level_type *level, const char *what)
{
  if (!g_comm) { fprintf(stderr, "hpgmg_b200: %s on a %d-rank level but no communicator was installed (hpgmg_b200_set_comm)\n", what, level->num_ranks); exit(1); }
  if (level->num_ranks != g_nranks) {
    fprintf(stderr, "hpgmg_b200: %s on a level shared by %d of %d ranks is not supported (sub-communicators, mg.c:979-997)\n", what, level->num_ranks, g_nranks);
    exit(1);
  }
}

extern "C" void hpgmg_comm_allreduce_slot_max(level_type *level, int slot)
{
  if (level->num_ranks <= 1 || g_nranks <= 1) return;
  require_world(level, "norm");
  double *s = hpgmg_rt_scalar_slots() + slot;
  NCCL_CHECK(N.AllReduce(s, s, 1, ncclFloat64, ncclMax, g_comm, g_stream));
  g_launches++;
}

static double allreduce_host_value(level_type *level, double v, int op, const char *what)
{
  if (level->num_ranks <= 1 || g_nranks <= 1) return v;
  require_world(level, what);
  const int slot = HPGMG_SLOT_SCRATCH + 3;
  double *s = hpgmg_rt_scalar_slots() + slot;
  CUDA_CHECK(cudaMemcpyAsync(s, &v, sizeof(double), cudaMemcpyHostToDevice, g_stream));
  CUDA_CHECK(cudaStreamSynchronize(g_stream));
  NCCL_CHECK(N.AllReduce(s, s, 1, ncclFloat64, op, g_comm, g_stream));
  double r = 0.0;
  hpgmg_rt_read_scalars(&r, slot, 1);
  return r;
}
extern "C" double hpgmg_comm_allreduce_max(level_type *level, double v) { return allreduce_host_value(level, v, ncclMax, "max-reduction"); }
extern "C" double hpgmg_comm_allreduce_sum(level_type *level, double v) { return allreduce_host_value(level, v, ncclSum, "sum-reduction"); }

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
