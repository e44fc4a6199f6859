// libppo_core.so — C ABI implementation (see include/ppo_core.h).  sm_100a only; no CPU fallback.
#include "../../include/ppo_core.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>

#include "device_common.cuh"
#include "host_rand.h"
#include "kernels_misc.cuh"
#include "kernels_mlp.cuh"
#include "kernels_mlp2.cuh"
#include "kernels_rollout.cuh"
#include "kernels_shuffle.cuh"
#include "kernels_small.cuh"
#include "kernels_umma.cuh"
#include "kernels_wide.cuh"
#include "meta_parser.h"

using namespace ppo;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(expr)                                                                                        \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) return fail(PPO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define TRY(expr)                 \
    do {                          \
        int _s = (expr);          \
        if (_s != PPO_OK) return _s; \
    } while (0)

extern "C" const char* ppo_last_error(void) { return g_err; }
extern "C" int ppo_abi_version(void) { return PPO_CORE_ABI_VERSION; }

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
namespace {
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueIdC { char internal[128]; };
enum { ncclSuccessC = 0 };
enum { ncclFloat32C = 7, ncclFloat64C = 8, ncclInt32C = 2 };
enum { ncclSumC = 0 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueIdC*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueIdC, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
#define LD(sym, field) field = reinterpret_cast<decltype(field)>(dlsym(lib, sym))
        LD("ncclGetUniqueId", GetUniqueId);
        LD("ncclCommInitRank", CommInitRank);
        LD("ncclCommDestroy", CommDestroy);
        LD("ncclAllReduce", AllReduce);
        LD("ncclAllGather", AllGather);
        LD("ncclGetErrorString", GetErrorString);
#undef LD
        return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && GetErrorString;
    }
};
NcclApi g_nccl;
}  // namespace

constexpr int F_TM_TRAIN = 64, F_NT_TRAIN = 512, F_TM_POLICY = 32, F_NT_POLICY = 256;

#include "abi_core.inl"
#include "abi_tensors.inl"
#include "abi_policy.inl"
#include "abi_comm.inl"
#include "abi_vecnorm.inl"
#include "abi_gae.inl"
#include "abi_rollout.inl"
#include "abi_update.inl"
#include "abi_introspect.inl"
