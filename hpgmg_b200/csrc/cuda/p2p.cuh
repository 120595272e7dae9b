/*
 * p2p.cuh -- device-side state and primitives of the direct peer-to-peer ghost exchange (see comm.cu for
 * the protocol and the setup, ghost.cu for the kernels).
 */
#ifndef HPGMG_B200_P2P_CUH
#define HPGMG_B200_P2P_CUH
#include "common.cuh"

#define P2P_MAX_NEIGHBOURS 32

struct P2PPlan {                                   /* device-resident state of one communicator */
  unsigned long long epoch_send, epoch_recv;       /* messages sent / received so far */
  unsigned int done_send, done_recv, late_passed;  /* thread blocks finished in the current kernel */
  unsigned int send_count[P2P_MAX_NEIGHBOURS], recv_count[P2P_MAX_NEIGHBOURS];
  int send_blocks[P2P_MAX_NEIGHBOURS], recv_blocks[P2P_MAX_NEIGHBOURS];   /* list entries per neighbour */
  unsigned long long *remote_data_flag[P2P_MAX_NEIGHBOURS];   /* in the receiver's arena: I write */
  unsigned long long *local_ack_flag[P2P_MAX_NEIGHBOURS];     /* in my arena: receiver writes, my pack waits */
  unsigned long long *local_data_flag[P2P_MAX_NEIGHBOURS];    /* in my arena: sender writes, my unpack waits */
  unsigned long long *remote_ack_flag[P2P_MAX_NEIGHBOURS];    /* in the sender's arena: I write */
};


__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

int hpgmg_comm_p2p_lookup(level_type *level, int shape, const blockCopy_type **pack, int *npack, const blockCopy_type **unpack, int *nunpack, P2PPlan **plan);

#endif
