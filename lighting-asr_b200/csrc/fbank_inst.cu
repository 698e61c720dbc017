// One translation unit per instantiation group of the fused kernel (see fbank_instances.h): compiled with
// -DB200FE_INST_GROUP=g by lighting-asr_b200/build.py, in parallel, then linked into libb200fe.so.
#include "fbank_kernel.cuh"
#include "fbank_instances.h"

#ifndef B200FE_INST_GROUP
#error "compile with -DB200FE_INST_GROUP=<0..B200FE_INST_GROUPS-1>"
#endif

#define B200FE_X(g, n, s, p, d, i, m, l, a) B200FE_X_##g(n, s, p, d, i, m, l, a)
#define B200FE_DEF(n, s, p, d, i, m, l, a) template __global__ void b200fe::fbank_fused_kernel<n, s, p, d, i, m, l, a>(const __grid_constant__ b200fe::FbankArgs);
#define B200FE_SKIP(n, s, p, d, i, m, l, a)
#define B200FE_X_0 B200FE_SKIP
#define B200FE_X_1 B200FE_SKIP
#define B200FE_X_2 B200FE_SKIP
#define B200FE_X_3 B200FE_SKIP
#define B200FE_X_4 B200FE_SKIP
#define B200FE_X_5 B200FE_SKIP
#define B200FE_X_6 B200FE_SKIP
#define B200FE_X_7 B200FE_SKIP
#if B200FE_INST_GROUP == 0
#undef B200FE_X_0
#define B200FE_X_0 B200FE_DEF
#elif B200FE_INST_GROUP == 1
#undef B200FE_X_1
#define B200FE_X_1 B200FE_DEF
#elif B200FE_INST_GROUP == 2
#undef B200FE_X_2
#define B200FE_X_2 B200FE_DEF
#elif B200FE_INST_GROUP == 3
#undef B200FE_X_3
#define B200FE_X_3 B200FE_DEF
#elif B200FE_INST_GROUP == 4
#undef B200FE_X_4
#define B200FE_X_4 B200FE_DEF
#elif B200FE_INST_GROUP == 5
#undef B200FE_X_5
#define B200FE_X_5 B200FE_DEF
#elif B200FE_INST_GROUP == 6
#undef B200FE_X_6
#define B200FE_X_6 B200FE_DEF
#elif B200FE_INST_GROUP == 7
#undef B200FE_X_7
#define B200FE_X_7 B200FE_DEF
#endif
B200FE_FBANK_INSTANCES(B200FE_X)

#if defined(B200FE_TIMELINE) && B200FE_INST_GROUP == 0
extern "C" int b200fe_debug_timeline(void* host_dst, unsigned long long bytes, int clear)
{
    if (clear) { void* p = nullptr; cudaGetSymbolAddress(&p, b200fe::g_timeline); return (int)cudaMemset(p, 0, sizeof(b200fe::g_timeline)); }
    if (bytes > sizeof(b200fe::g_timeline)) bytes = sizeof(b200fe::g_timeline);
    return (int)cudaMemcpyFromSymbol(host_dst, b200fe::g_timeline, bytes);
}
#endif
