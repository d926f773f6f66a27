// lat <-> m exchange of the spatially sharded transform over NVLink peer memory (BASELINE config 5).
//
// One process per GPU.  Every rank allocates its exchange buffers with cudaMalloc, exports them through CUDA IPC and maps
// every peer's buffers (msfno_peer_alloc / msfno_peer_open); the transpose between the longitude stage (latitudes sharded)
// and the Legendre stage (orders sharded) is then ONE kernel per direction that stores each destination's sub-block
// straight into that destination's operand buffer through NVSwitch (msfno_peer_block_copy) -- no staging copy on either
// side and no library collective -- bracketed by a flag barrier in the same peer memory (msfno_peer_barrier).  Measured on
// 2 x B200: a 256 MB peer store runs at 758 GB/s per direction (tools/microbench/ipc_probe.py), where
// pack -> ncclSend/Recv -> unpack moved the same bytes in ~1.3 ms per 177 MB.
//
// The reference has no distributed transform (DDP only: /root/reference main.py:39-49, MSFNO/Models/train.py:370-374).
#include <string.h>

#include "common.cuh"

using namespace msfno;

namespace {

struct PeerFlags {
  unsigned int* flags[MSFNO_MAX_PEERS];   // flags[r]: rank r's flag array [world] (this rank's mapping of it)
};

struct BlockSet {
  int n;
  msfno_peer_block b[MSFNO_MAX_PEERS];
};

// grid = (row chunks, blocks): warp per row, lanes along the latitude-contiguous columns.  VEC: every offset, pitch and
// count is a multiple of four floats and both bases are 16-byte aligned (the latitude split of DistributedSHT is aligned
// to four rows for this) -- 16-byte loads and NVLink stores.
template <bool VEC>
__global__ void peer_block_copy_kernel(const float* __restrict__ src, BlockSet bs) {
  const msfno_peer_block& k = bs.b[blockIdx.y];
  if (!k.dst || k.rows <= 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int ncopy = k.cols, nfill = k.cols + k.zero_tail;
  for (long long r = (long long)blockIdx.x * nw + warp; r < k.rows; r += (long long)gridDim.x * nw) {
    const float* s = src + (k.src_row0 + r) * k.src_pitch + k.src_col0;
    float* d = k.dst + (k.dst_row0 + r) * k.dst_pitch + k.dst_col0;
    if (VEC) {
      const float4* s4 = reinterpret_cast<const float4*>(s);
      float4* d4 = reinterpret_cast<float4*>(d);
      for (int j = lane; j < (nfill >> 2); j += 32) d4[j] = (4 * j < ncopy) ? s4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int j = lane; j < nfill; j += 32) d[j] = j < ncopy ? s[j] : 0.0f;
    }
  }
}

// One thread per peer: publish `epoch` in the peer's flag slot of this rank (release at system scope: everything this
// GPU wrote before -- including the stores of earlier kernels in the stream -- is visible to a peer that has seen the
// flag), then wait until that peer has published the same epoch here.  The spin is bounded: a peer that never arrives
// (crashed process) raises an error flag instead of hanging the GPU.
// The epoch lives in device memory (state[1], incremented here): the launch carries no per-call host value, so a
// captured CUDA graph of the exchange replays correctly.
__global__ void peer_barrier_kernel(PeerFlags pf, int rank, int world, unsigned int* state) {
  __shared__ unsigned int epoch_s;
  if (threadIdx.x == 0) epoch_s = ++state[1];
  __syncthreads();
  const unsigned int epoch = epoch_s;
  int* timed_out = reinterpret_cast<int*>(state);
  const int r = threadIdx.x;
  if (r >= world) return;
  __threadfence_system();
  unsigned int* theirs = pf.flags[r] + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(theirs), "r"(epoch) : "memory");
  const unsigned int* mine = pf.flags[rank] + r;
  for (int it = 0; it < (1 << 22); ++it) {   // ~ seconds: local acquire loads + a short sleep every 1024 polls
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(mine) : "memory");
    if ((int)(v - epoch) >= 0) return;
    if ((it & 1023) == 1023) __nanosleep(200);
  }
  *timed_out = 1;
}

}  // namespace

extern "C" {

int msfno_peer_alloc(size_t bytes, void** ptr, void* handle64) {
  if (!ptr || !handle64 || bytes == 0) return record_error(MSFNO_ERR_BAD_SHAPE, "peer_alloc: bad argument");
  void* p = nullptr;
  MSFNO_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return record_cuda_error(e, __FILE__, __LINE__);
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  *ptr = p;
  return MSFNO_OK;
}

int msfno_peer_free(void* ptr) {
  if (!ptr) return MSFNO_OK;
  MSFNO_CUDA_OK(cudaFree(ptr));
  return MSFNO_OK;
}

int msfno_peer_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return record_error(MSFNO_ERR_BAD_SHAPE, "peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  MSFNO_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MSFNO_OK;
}

int msfno_peer_close(void* ptr) {
  if (!ptr) return MSFNO_OK;
  MSFNO_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return MSFNO_OK;
}

int msfno_peer_block_copy(const float* src, int nblocks, const msfno_peer_block* blocks, void* stream) {
  if (!src || !blocks || nblocks < 1 || nblocks > MSFNO_MAX_PEERS) return record_error(MSFNO_ERR_BAD_SHAPE, "peer_block_copy: bad argument");
  BlockSet bs{};
  bs.n = nblocks;
  long long maxrows = 0;
  bool vec = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (int i = 0; i < nblocks; ++i) {
    const msfno_peer_block& k = blocks[i];
    if (k.rows < 0 || k.cols < 0 || k.zero_tail < 0 || k.src_pitch < k.src_col0 + k.cols || k.dst_pitch < k.dst_col0 + k.cols + k.zero_tail ||
        (k.rows > 0 && !k.dst))
      return record_error(MSFNO_ERR_BAD_SHAPE, "peer_block_copy: bad block");
    bs.b[i] = k;
    if (k.rows > maxrows) maxrows = k.rows;
    if (k.rows > 0 && ((reinterpret_cast<uintptr_t>(k.dst) & 15) || ((k.cols | k.zero_tail | k.src_col0 | k.src_pitch | k.dst_col0 | k.dst_pitch) & 3)))
      vec = false;
  }
  if (maxrows == 0) return MSFNO_OK;
  long long chunks = (maxrows + 7) / 8;
  if (chunks > 148 * 4) chunks = 148 * 4;
  if (vec) peer_block_copy_kernel<true><<<dim3((unsigned)chunks, nblocks), 256, 0, (cudaStream_t)stream>>>(src, bs);
  else peer_block_copy_kernel<false><<<dim3((unsigned)chunks, nblocks), 256, 0, (cudaStream_t)stream>>>(src, bs);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

int msfno_peer_barrier(unsigned int* const* flags, int rank, int world, unsigned int* state, void* stream) {
  if (!flags || !state || world < 1 || world > MSFNO_MAX_PEERS || rank < 0 || rank >= world)
    return record_error(MSFNO_ERR_BAD_SHAPE, "peer_barrier: bad argument");
  PeerFlags pf{};
  for (int r = 0; r < world; ++r) {
    if (!flags[r]) return record_error(MSFNO_ERR_BAD_SHAPE, "peer_barrier: missing flag buffer");
    pf.flags[r] = flags[r];
  }
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pf, rank, world, state);
  count_launch();
  MSFNO_CUDA_OK(cudaGetLastError());
  return MSFNO_OK;
}

}  // extern "C"
