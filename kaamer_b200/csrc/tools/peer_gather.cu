// peer_gather.cu — how fast can one B200 read random 8-byte table entries from a PEER GPU's HBM
// through NVLink (mode P)?  Variants: load flavour, bytes per probe, table size (TLB reach),
// mapping kind (same-process peer access vs cross-process CUDA IPC), local/remote mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_gather peer_gather.cu -lcuda && ./peer_gather
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <unistd.h>
#include <sys/wait.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
template <int V>
__device__ __forceinline__ uint64_t ld(const uint64_t *p) {
  uint64_t v = 0, w;
  if (V == 0) asm volatile("ld.global.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 1) asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 3) asm volatile("ld.global.cg.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 4) asm volatile("ld.volatile.global.b64 %0, [%1];" : "=l"(v) : "l"(p));
  if (V == 5) { asm volatile("ld.global.v2.b64 {%0,%1}, [%2];" : "=l"(v), "=l"(w) : "l"((const uint64_t *)((uintptr_t)p & ~(uintptr_t)15))); v += w; }
  if (V == 6) asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
// remote_per_256: how many of 256 probes go to the remote table (the rest to the local one)
template <int V, int U>
__global__ void k_gather(const uint64_t *__restrict__ remote, const uint64_t *__restrict__ local, uint64_t slots,
                         uint64_t n, uint64_t *out, uint64_t seed, unsigned remote_per_256) {
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (uint64_t i = tid; i < n; i += stride * U) {
    uint64_t v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint64_t j = i + (uint64_t)u * stride;
      uint64_t h = mix(j + seed);
      uint64_t idx = h % slots;
      const uint64_t *t = ((h >> 40) & 255u) < remote_per_256 ? remote : local;
      v[u] = j < n ? ld<V>(t + idx) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
  }
  if (acc == 0x1234567) out[0] = acc;
}
__global__ void k_stream(const uint4 *__restrict__ src, uint64_t n16, uint64_t *out) {
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (uint64_t i = tid; i < n16; i += stride) { uint4 v = src[i]; acc += v.x + v.y + v.z + v.w; }
  if (acc == 0x1234567) out[0] = acc;
}
template <int V, int U>
void run(const char *name, const uint64_t *remote, const uint64_t *local, uint64_t slots, uint64_t n, uint64_t *out,
         unsigned rp256, int ctas_per_sm = 8) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    k_gather<V, U><<<148 * ctas_per_sm, 256>>>(remote, local, slots, n, out, rep * 7919 + 1, rp256);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  printf("  %-52s remote %3u/256  %8.3f ms  %7.3f G probes/s  %s\n", name, rp256, best, n / best / 1e6,
         cudaGetErrorString(cudaGetLastError()));
  fflush(stdout);
}
static void suite(const char *what, const uint64_t *remote, const uint64_t *local, uint64_t slots, uint64_t *out) {
  printf("== %s, %.2f GB table ==\n", what, slots * 8 / 1e9);
  const uint64_t n = 1ull << 24;
  run<0, 4>("ld.global x4", remote, local, slots, n, out, 256);
  run<1, 4>("ld.global.nc x4", remote, local, slots, n, out, 256);
  run<2, 4>("ld.global.nc.L1::no_allocate.L2::64B x4", remote, local, slots, n, out, 256);
  run<3, 4>("ld.global.cg x4", remote, local, slots, n, out, 256);
  run<4, 4>("ld.volatile.global x4", remote, local, slots, n, out, 256);
  run<5, 4>("ld.global.v2.b64 (16 B) x4", remote, local, slots, n, out, 256);
  run<6, 4>("ld.relaxed.sys x4", remote, local, slots, n, out, 256);
  run<0, 1>("ld.global x1", remote, local, slots, n, out, 256);
  run<0, 8>("ld.global x8", remote, local, slots, n, out, 256);
  run<0, 4>("ld.global x4, 2 CTAs/SM", remote, local, slots, n, out, 256, 2);
  run<0, 4>("ld.global x4 mixed", remote, local, slots, n * 4, out, 128);
  run<0, 4>("ld.global x4 mixed", remote, local, slots, n * 4, out, 32);
  run<0, 4>("ld.global x4 local only", remote, local, slots, n * 4, out, 0);
}
#define CD(x) do { CUresult r = (x); if (r != CUDA_SUCCESS) { const char *m = 0; cuGetErrorString(r, &m); printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, m ? m : "?"); exit(1);} } while (0)
static void send_fd(int sock, int fd) {
  char b = 'x'; struct iovec io = {&b, 1}; char ctl[CMSG_SPACE(sizeof(int))]; memset(ctl, 0, sizeof ctl);
  struct msghdr m = {}; m.msg_iov = &io; m.msg_iovlen = 1; m.msg_control = ctl; m.msg_controllen = sizeof ctl;
  struct cmsghdr *c = CMSG_FIRSTHDR(&m); c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  if (sendmsg(sock, &m, 0) != 1) { printf("sendmsg failed\n"); exit(1); }
}
static int recv_fd(int sock) {
  char b; struct iovec io = {&b, 1}; char ctl[CMSG_SPACE(sizeof(int))];
  struct msghdr m = {}; m.msg_iov = &io; m.msg_iovlen = 1; m.msg_control = ctl; m.msg_controllen = sizeof ctl;
  if (recvmsg(sock, &m, 0) != 1) { printf("recvmsg failed\n"); exit(1); }
  int fd = -1; memcpy(&fd, CMSG_DATA(CMSG_FIRSTHDR(&m)), sizeof(int));
  return fd;
}
static CUmemAllocationProp vmm_prop(int dev) {
  CUmemAllocationProp p = {}; p.type = CU_MEM_ALLOCATION_TYPE_PINNED; p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE; p.location.id = dev; return p;
}
static CUdeviceptr vmm_map(CUmemGenericAllocationHandle h, size_t bytes, int dev, size_t align = 2 << 20) {
  CUdeviceptr va; CD(cuMemAddressReserve(&va, bytes, align, 0, 0)); CD(cuMemMap(va, bytes, 0, h, 0));
  CUmemAccessDesc a = {}; a.location.type = CU_MEM_LOCATION_TYPE_DEVICE; a.location.id = dev; a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CD(cuMemSetAccess(va, bytes, &a, 1)); return va;
}
static void mini(const char *what, const uint64_t *remote, const uint64_t *local, uint64_t slots, uint64_t *out) {
  printf("== %s, %.2f GB table ==\n", what, slots * 8 / 1e9);
  const uint64_t n = 1ull << 24;
  run<0, 4>("ld.global x4", remote, local, slots, n, out, 256);
  run<0, 4>("ld.global x4 mixed", remote, local, slots, n * 4, out, 128);
}
int main(int argc, char **argv) {
  int nd = 0; 
  const uint64_t slots_big = 1813366968ull / 2, slots_small = (256ull << 20) / 8;
  // ---- cross-process part first (fork before any CUDA call in this process) ----
  int p2c[2], c2p[2], sp[2];
  if (pipe(p2c) || pipe(c2p) || socketpair(AF_UNIX, SOCK_STREAM, 0, sp)) return 1;
  const size_t vbytes = (slots_big * 8 + (2 << 20) - 1) / (2 << 20) * (2 << 20);
  const size_t hbytes = (slots_big * 8 + ((size_t)512 << 20) - 1) / ((size_t)512 << 20) * ((size_t)512 << 20);
  pid_t child2 = fork();
  if (child2 == 0) {  // exporter of a VMM allocation on device 1
    CK(cudaSetDevice(1)); CK(cudaFree(0));
    CUmemAllocationProp p = vmm_prop(1); CUmemGenericAllocationHandle h; CD(cuMemCreate(&h, vbytes, &p, 0));
    CUdeviceptr va = vmm_map(h, vbytes, 1);
    CK(cudaMemset((void *)va, 1, vbytes)); CK(cudaDeviceSynchronize());
    int fd; CD(cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    send_fd(sp[1], fd);
    {
      size_t gmin = 0, grec = 0;
      cuMemGetAllocationGranularity(&gmin, &p, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
      cuMemGetAllocationGranularity(&grec, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
      printf("VMM granularity: minimum %zu, recommended %zu\n", gmin, grec); fflush(stdout);
    }
    CUmemGenericAllocationHandle h2; CD(cuMemCreate(&h2, hbytes, &p, 0));
    CUdeviceptr va2 = vmm_map(h2, hbytes, 1, (size_t)512 << 20);
    CK(cudaMemset((void *)va2, 1, hbytes)); CK(cudaDeviceSynchronize());
    int fd2; CD(cuMemExportToShareableHandle(&fd2, h2, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    send_fd(sp[1], fd2);
    char c; if (read(sp[1], &c, 1) != 1) return 1;
    return 0;
  }
  pid_t child = fork();
  if (child == 0) {
    CK(cudaSetDevice(1));
    uint64_t *t; CK(cudaMalloc(&t, slots_big * 8)); CK(cudaMemset(t, 1, slots_big * 8)); CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, t));
    if (write(c2p[1], &h, sizeof h) != sizeof h) return 1;
    char c; if (read(p2c[0], &c, 1) != 1) return 1;  // wait until the parent is done
    cudaFree(t);
    return 0;
  }
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 1; }
  CK(cudaSetDevice(0));
  int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1)); printf("canAccessPeer(0,1) = %d\n", can);
  int attr = 0;
  cudaDeviceGetP2PAttribute(&attr, cudaDevP2PAttrPerformanceRank, 0, 1); printf("p2p performance rank = %d\n", attr);
  cudaDeviceGetP2PAttribute(&attr, cudaDevP2PAttrNativeAtomicSupported, 0, 1); printf("p2p native atomics = %d\n", attr);
  uint64_t *local, *out;
  CK(cudaMalloc(&local, slots_big * 8)); CK(cudaMemset(local, 1, slots_big * 8)); CK(cudaMalloc(&out, 8));
  {
    cudaIpcMemHandle_t h;
    if (read(c2p[0], &h, sizeof h) != sizeof h) { printf("child failed\n"); return 1; }
    void *r = nullptr;
    CK(cudaIpcOpenMemHandle(&r, h, cudaIpcMemLazyEnablePeerAccess));
    if (argc > 1) {
      mini("cross-process CUDA IPC mapping", (const uint64_t *)r, local, slots_big, out);
    } else {
      suite("cross-process CUDA IPC mapping", (const uint64_t *)r, local, slots_big, out);
      suite("cross-process CUDA IPC mapping (first 256 MB only)", (const uint64_t *)r, local, slots_small, out);
    }
    // streaming read of the remote table
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a); k_stream<<<148 * 8, 512>>>((const uint4 *)r, (1ull << 30) / 16, out); cudaEventRecord(b);
      cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep) printf("  streaming read of 1 GiB remote: %.3f ms = %.1f GB/s\n", ms, (1ull << 30) / ms / 1e6);
    }
    CK(cudaIpcCloseMemHandle(r));
    char c = 1; if (write(p2c[1], &c, 1) != 1) return 1;
    int st; waitpid(child, &st, 0);
  }
  // ---- cross-process CUDA VMM (cuMemCreate / POSIX fd / cuMemMap): what libkaamer_gpu uses ----
  {
    int fd = recv_fd(sp[0]);
    CUmemGenericAllocationHandle h; CD(cuMemImportFromShareableHandle(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUdeviceptr va = vmm_map(h, vbytes, 0);
    suite("cross-process CUDA VMM mapping (posix fd, 2 MiB pages)", (const uint64_t *)va, local, slots_big, out);
    mini("cross-process CUDA VMM mapping (first 256 MB only)", (const uint64_t *)va, local, slots_small, out);
    for (int mb : {512, 1024, 1536, 2048, 3072, 4096, 6144}) {
      char nm[96]; snprintf(nm, sizeof nm, "cross-process CUDA VMM mapping (first %d MB only)", mb);
      mini(nm, (const uint64_t *)va, local, ((uint64_t)mb << 20) / 8, out);
    }
    CD(cuMemUnmap(va, vbytes)); CD(cuMemRelease(h)); CD(cuMemAddressFree(va, vbytes));
    int fd2 = recv_fd(sp[0]);
    CUmemGenericAllocationHandle h2; CD(cuMemImportFromShareableHandle(&h2, (void *)(uintptr_t)fd2, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUdeviceptr va2 = vmm_map(h2, hbytes, 0, (size_t)512 << 20);
    mini("cross-process CUDA VMM mapping, size and VA aligned to 512 MiB", (const uint64_t *)va2, local, slots_big, out);
    CD(cuMemUnmap(va2, hbytes)); CD(cuMemRelease(h2)); CD(cuMemAddressFree(va2, hbytes));
    char c = 1; if (write(sp[0], &c, 1) != 1) return 1;
    int st; waitpid(child2, &st, 0);
  }
  // ---- same-process peer access ----
  {
    CK(cudaSetDevice(1));
    uint64_t *remote; CK(cudaMalloc(&remote, slots_big * 8)); CK(cudaMemset(remote, 1, slots_big * 8)); CK(cudaDeviceSynchronize());
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    suite("same-process cudaDeviceEnablePeerAccess", remote, local, slots_big, out);
    suite("same-process peer access (first 256 MB only)", remote, local, slots_small, out);
  }
  {  // local VMM allocation (2 MiB pages) against the cudaMalloc'ed local table
    CUmemAllocationProp p = vmm_prop(0); CUmemGenericAllocationHandle h; CD(cuMemCreate(&h, vbytes, &p, 0));
    CUdeviceptr va = vmm_map(h, vbytes, 0);
    CK(cudaMemset((void *)va, 1, vbytes)); CK(cudaDeviceSynchronize());
    printf("== LOCAL table in a CUDA VMM allocation (2 MiB pages), %.2f GB ==\n", slots_big * 8 / 1e9);
    run<0, 4>("ld.global x4 local VMM", (const uint64_t *)va, (const uint64_t *)va, slots_big, 1ull << 26, out, 0);
    run<0, 4>("ld.global x4 local cudaMalloc", local, local, slots_big, 1ull << 26, out, 0);
    CD(cuMemUnmap(va, vbytes)); CD(cuMemRelease(h)); CD(cuMemAddressFree(va, vbytes));
  }
  {  // same-process VMM allocation on device 1 with access granted to device 0
    CUmemAllocationProp p = vmm_prop(1); CUmemGenericAllocationHandle h; CD(cuMemCreate(&h, hbytes, &p, 0));
    CUdeviceptr va = vmm_map(h, hbytes, 1, (size_t)512 << 20);
    CUmemAccessDesc a = {}; a.location.type = CU_MEM_LOCATION_TYPE_DEVICE; a.location.id = 0; a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CD(cuMemSetAccess(va, hbytes, &a, 1));
    CK(cudaSetDevice(1)); CK(cudaMemset((void *)va, 1, hbytes)); CK(cudaDeviceSynchronize()); CK(cudaSetDevice(0));
    mini("same-process CUDA VMM allocation, access granted to device 0 (512 MiB aligned)", (const uint64_t *)va, local, slots_big, out);
  }
  printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
