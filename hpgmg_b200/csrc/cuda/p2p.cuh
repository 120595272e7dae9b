/*
 * p2p.cuh -- device-side state and primitives of the direct peer-to-peer ghost exchange (see comm.cu for
 * the setup, ghost.cu for the kernel).
 *
 * Protocol ("LL", the low-latency scheme NCCL uses for small messages): a message element is one
 * 16-byte slot {lo32(data), flag, hi32(data), flag} written with a single 128-bit volatile store into
 * the RECEIVER's memory over NVLink.  The flag is the exchange's sequence number, so the receiver
 * simply polls the slot until both flags match -- data and "ready" travel in the same 8-byte-atomic
 * units: no fences, no separate flag, no acknowledgement.  Each neighbour pair has TWO slot arrays
 * used alternately (sequence parity).  Overwriting array p at exchange k+2 is safe because ghost
 * exchange is symmetric: I only get there after unpacking the neighbour's message k+1, which it sent
 * after it had unpacked my message k.
 */
#ifndef HPGMG_B200_P2P_CUH
#define HPGMG_B200_P2P_CUH
#include "common.cuh"

#define P2P_MAX_NEIGHBOURS 32

struct P2PPlan {                                   /* device-resident state of one communicator */
  unsigned long long epoch;                        /* exchanges completed so far */
  unsigned int done, unpack_done;                  /* thread blocks finished in the current launch */
  int send_size[P2P_MAX_NEIGHBOURS], recv_size[P2P_MAX_NEIGHBOURS];   /* doubles per message */
  uint4 *ll_remote[P2P_MAX_NEIGHBOURS];            /* 2 x send_size slots in the receiver's arena: I write */
  uint4 *ll_local[P2P_MAX_NEIGHBOURS];             /* 2 x recv_size slots in my arena: the sender writes, I poll */
};

__device__ __forceinline__ void ll_store(uint4 *slot, const double v, const unsigned flag)
{
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
__device__ __forceinline__ double ll_load(const uint4 *slot, const unsigned flag)
{
  unsigned lo, f0, hi, f1;
  do {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(slot) : "memory");
  } while (f0 != flag || f1 != flag);
  return __hiloint2double((int)hi, (int)lo);
}

int hpgmg_comm_p2p_lookup(level_type *level, int shape, const blockCopy_type **pack, int *npack, const blockCopy_type **unpack, int *nunpack, P2PPlan **plan);

#endif
