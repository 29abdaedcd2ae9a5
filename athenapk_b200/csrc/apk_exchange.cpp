// apk_exchange.cpp -- C-ABI of the multi-rank halo exchange (include/apk_b200.h: apk_halo_*,
// apk_p2p_*, apk_stage_barrier, apk_allreduce_scalar).  Host C++ over NCCL and the CUDA runtime;
// the pack / unpack kernels live in apk_halo.cu.  Replaces the Send/Receive/SetBounds machinery
// of external/parthenon/src/bvals/comms/boundary_communication.cpp:48-348 and the per-buffer
// messages of utils/communication_buffer.hpp:209-236 for packs whose blocks span several GPUs.
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "apk_launch.h"

struct apk_halo_plan {
  std::vector<apk_halo_peer_t> peers;
  int nsend = 0, nrecv = 0;
  apk_halo_item_t *d_send_items = nullptr, *d_recv_items = nullptr;
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
  ncclComm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t packed = nullptr, done = nullptr;
  bool inflight = false;
};

namespace {
inline cudaStream_t S(void *s) { return reinterpret_cast<cudaStream_t>(s); }
int fail(const char *what, cudaError_t e) {
  static thread_local char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  apk_common::set_error(buf);
  return 2;
}
int nccl_fail(const char *what, ncclResult_t r) {
  static thread_local char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, ncclGetErrorString(r));
  apk_common::set_error(buf);
  return 3;
}
#define CU(call, what)                         \
  do {                                         \
    cudaError_t e_ = (call);                   \
    if (e_ != cudaSuccess) return fail(what, e_); \
  } while (0)
}  // namespace

extern "C" {

int apk_halo_plan_create(const apk_halo_item_t *h_send, int nsend, const apk_halo_item_t *h_recv,
                         int nrecv, const apk_halo_peer_t *h_peers, int npeers,
                         long long send_total, long long recv_total, void *nccl_comm,
                         apk_halo_plan_t **out) {
  if (!out || nsend < 0 || nrecv < 0 || npeers < 0 || (npeers > 0 && !h_peers) || !nccl_comm) {
    apk_common::set_error("apk_halo_plan_create: bad arguments");
    return 1;
  }
  apk_halo_plan *p = new apk_halo_plan();
  p->peers.assign(h_peers, h_peers + npeers);
  p->nsend = nsend;
  p->nrecv = nrecv;
  p->comm = reinterpret_cast<ncclComm_t>(nccl_comm);
  const size_t one = 1;
  CU(cudaMalloc(&p->d_sendbuf, std::max<size_t>((size_t)send_total, one) * sizeof(double)), "sendbuf");
  CU(cudaMalloc(&p->d_recvbuf, std::max<size_t>((size_t)recv_total, one) * sizeof(double)), "recvbuf");
  CU(cudaMalloc(&p->d_send_items, std::max<size_t>(nsend, one) * sizeof(apk_halo_item_t)), "items");
  CU(cudaMalloc(&p->d_recv_items, std::max<size_t>(nrecv, one) * sizeof(apk_halo_item_t)), "items");
  if (nsend)
    CU(cudaMemcpy(p->d_send_items, h_send, nsend * sizeof(apk_halo_item_t), cudaMemcpyHostToDevice),
       "items");
  if (nrecv)
    CU(cudaMemcpy(p->d_recv_items, h_recv, nrecv * sizeof(apk_halo_item_t), cudaMemcpyHostToDevice),
       "items");
  CU(cudaStreamCreateWithFlags(&p->comm_stream, cudaStreamNonBlocking), "comm stream");
  CU(cudaEventCreateWithFlags(&p->packed, cudaEventDisableTiming), "event");
  CU(cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming), "event");
  *out = p;
  return 0;
}

void apk_halo_plan_destroy(apk_halo_plan_t *p) {
  if (!p) return;
  if (p->comm_stream) cudaStreamSynchronize(p->comm_stream);
  cudaFree(p->d_sendbuf);
  cudaFree(p->d_recvbuf);
  cudaFree(p->d_send_items);
  cudaFree(p->d_recv_items);
  if (p->packed) cudaEventDestroy(p->packed);
  if (p->done) cudaEventDestroy(p->done);
  if (p->comm_stream) cudaStreamDestroy(p->comm_stream);
  delete p;
}

int apk_halo_start(apk_halo_plan_t *p, const apk_pack_desc_t *d, const apk_block_t *blocks,
                   int with_prim, void *stream) {
  if (!p || !d || !blocks) {
    apk_common::set_error("apk_halo_start: null plan / descriptor / blocks");
    return 1;
  }
  apk_common::launch_halo_pack(*d, blocks, p->d_send_items, p->nsend, with_prim, p->d_sendbuf,
                               S(stream));
  CU(cudaEventRecord(p->packed, S(stream)), "record");
  CU(cudaStreamWaitEvent(p->comm_stream, p->packed, 0), "wait");
  // offsets / counts are sized for cons+prim; a single-field exchange packs densely at half of
  // every offset (all slab sizes are even multiples), so only live bytes travel
  const int div = (with_prim == 1) ? 1 : 2;
  ncclGroupStart();
  for (const apk_halo_peer_t &q : p->peers) {
    ncclSend(p->d_sendbuf + q.send_off / div, (size_t)(q.send_cnt / div), ncclDouble, q.rank, p->comm,
             p->comm_stream);
    ncclRecv(p->d_recvbuf + q.recv_off / div, (size_t)(q.recv_cnt / div), ncclDouble, q.rank, p->comm,
             p->comm_stream);
  }
  ncclResult_t r = ncclGroupEnd();
  if (r != ncclSuccess) return nccl_fail("apk_halo_start", r);
  CU(cudaEventRecord(p->done, p->comm_stream), "record");
  p->inflight = true;
  return 0;
}

int apk_halo_test(apk_halo_plan_t *p) {
  if (!p || !p->inflight) return 0;
  cudaError_t e = cudaEventQuery(p->done);
  if (e == cudaSuccess) return 0;
  if (e == cudaErrorNotReady) return 1;
  return fail("apk_halo_test", e);
}

int apk_halo_finish(apk_halo_plan_t *p, const apk_pack_desc_t *d, const apk_block_t *blocks,
                    int with_prim, void *stream) {
  if (!p || !d || !blocks) {
    apk_common::set_error("apk_halo_finish: null plan / descriptor / blocks");
    return 1;
  }
  if (!p->inflight) return 0;
  CU(cudaStreamWaitEvent(S(stream), p->done, 0), "wait comm");
  apk_common::launch_halo_unpack(*d, blocks, p->d_recv_items, p->nrecv, with_prim, p->d_recvbuf,
                                 S(stream));
  p->inflight = false;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail("apk_halo_finish", e);
}

int apk_p2p_map_peers(void *nccl_comm, int rank, int nranks, void *const *h_local, int nbufs,
                      const int *h_need, void **h_mapped, int *all_ok, void *stream) {
  if (!nccl_comm || !h_local || !h_need || !h_mapped || !all_ok || nbufs < 1 || nbufs > 8 ||
      nranks < 1) {
    apk_common::set_error("apk_p2p_map_peers: bad arguments");
    return 1;
  }
  ncclComm_t c = reinterpret_cast<ncclComm_t>(nccl_comm);
  cudaStream_t st = S(stream);
  struct Handles {
    cudaIpcMemHandle_t h[8];
  };
  int ok = 1;
  Handles mine{};
  for (int f = 0; f < nbufs; ++f)
    if (cudaIpcGetMemHandle(&mine.h[f], h_local[f]) != cudaSuccess) ok = 0;
  cudaGetLastError();
  std::vector<Handles> all(nranks);
  char *d_all = nullptr;
  int *d_flag = nullptr;
  CU(cudaMalloc(&d_all, sizeof(Handles) * nranks), "ipc gather");
  CU(cudaMalloc(&d_flag, sizeof(int)), "ipc flag");
  CU(cudaMemcpyAsync(d_all + sizeof(Handles) * rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice,
                     st), "ipc handles");
  ncclAllGather(d_all + sizeof(Handles) * rank, d_all, sizeof(Handles), ncclChar, c, st);
  CU(cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * nranks, cudaMemcpyDeviceToHost, st),
     "ipc handles");
  CU(cudaStreamSynchronize(st), "ipc handles");
  for (int i = 0; i < nranks * nbufs; ++i) h_mapped[i] = nullptr;
  for (int r = 0; r < nranks && ok; ++r) {
    if (r == rank || !h_need[r]) continue;
    for (int f = 0; f < nbufs; ++f) {
      void *m = nullptr;
      if (cudaIpcOpenMemHandle(&m, all[r].h[f], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ok = 0;
        cudaGetLastError();
        break;
      }
      h_mapped[r * nbufs + f] = m;
    }
  }
  // all ranks must agree
  CU(cudaMemcpyAsync(d_flag, &ok, sizeof(int), cudaMemcpyHostToDevice, st), "p2p flag");
  ncclAllReduce(d_flag, d_flag, 1, ncclInt, ncclMin, c, st);
  int agreed = 0;
  CU(cudaMemcpyAsync(&agreed, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st), "p2p flag");
  CU(cudaStreamSynchronize(st), "p2p flag");
  cudaFree(d_all);
  cudaFree(d_flag);
  if (!agreed) apk_p2p_unmap(h_mapped, nranks * nbufs);
  *all_ok = agreed;
  return 0;
}

int apk_p2p_unmap(void **h_mapped, int n) {
  for (int i = 0; i < n; ++i) {
    if (h_mapped[i]) cudaIpcCloseMemHandle(h_mapped[i]);
    h_mapped[i] = nullptr;
  }
  cudaGetLastError();
  return 0;
}

int apk_stage_barrier(void *nccl_comm, int *d_word, void *stream) {
  ncclResult_t r = ncclAllReduce(d_word, d_word, 1, ncclInt, ncclSum,
                                 reinterpret_cast<ncclComm_t>(nccl_comm), S(stream));
  return r == ncclSuccess ? 0 : nccl_fail("apk_stage_barrier", r);
}

int apk_allreduce_scalar(void *nccl_comm, void *d_value, int op, void *stream) {
  ncclComm_t c = reinterpret_cast<ncclComm_t>(nccl_comm);
  ncclResult_t r = (op == 0) ? ncclAllReduce(d_value, d_value, 1, ncclDouble, ncclMin, c, S(stream))
                             : ncclAllReduce(d_value, d_value, 1, ncclUint64, ncclSum, c, S(stream));
  return r == ncclSuccess ? 0 : nccl_fail("apk_allreduce_scalar", r);
}

}  // extern "C"
