// sdf_jit_entry.cuh -- kernel entry points of the run-time compiled translation unit.
// Included AFTER the generated `struct SpecEval` (sdfb200.cu: jit_source); SDF_JIT_KERNEL selects
// the one kernel this compilation emits, always under the name `jit_kernel`.
#pragma once

#if SDF_JIT_KERNEL == 0
extern "C" __global__ void __launch_bounds__(128)
jit_kernel(const GridDesc g, int sparse, unsigned char* flags, unsigned int* active_list, unsigned int* batch_tris,
           Counters* ctr) {
    skip_prepass_body(SpecEval(), g, sparse, flags, active_list, batch_tris, ctr);
}
#elif SDF_JIT_KERNEL == 4          // precision `exact`: sign masks by ballot, no value tile
extern "C" __global__ void __launch_bounds__(SIGN_THREADS, SIGN_BLOCKS_PER_SM)
jit_kernel(const GridDesc g, const float tau, const unsigned int* active_list, const unsigned int work_begin,
           const unsigned int work_end, Counters* ctr, unsigned char* scratch) {
    tile_sign_body<32, SIGN_THREADS>(SpecEval(), g, tau, active_list, work_begin, work_end, ctr, scratch);
}
#elif SDF_JIT_KERNEL >= 1 && SDF_JIT_KERNEL <= 3
#define SDF_JIT_TILE_THREADS ((SDF_JIT_KERNEL - 1) == SDFB200_PREC_F64 ? 512 : 1024)
extern "C" __global__ void __launch_bounds__(SDF_JIT_TILE_THREADS, 1)
jit_kernel(const GridDesc g, const float tau, const unsigned int* active_list, const unsigned int work_begin,
           const unsigned int work_end, const TileLists L, Counters* ctr) {
    tile_eval_body<32, SDF_JIT_KERNEL - 1, SDF_JIT_TILE_THREADS>(SpecEval(), g, tau, active_list, work_begin, work_end, L, ctr);
}
#elif SDF_JIT_KERNEL == 8
extern "C" __global__ void __launch_bounds__(128)
jit_kernel(const GridDesc g, const TileLists L, Counters* ctr) { refine_points_body<32>(SpecEval(), g, L, ctr); }
#elif SDF_JIT_KERNEL == 5
extern "C" __global__ void __launch_bounds__(STREAM_THREADS)
jit_kernel(const float* xyz, long long n, float* out) { eval_stream_f32_body<false>(SpecEval(), xyz, n, out); }
#elif SDF_JIT_KERNEL == 6
extern "C" __global__ void __launch_bounds__(STREAM_THREADS)
jit_kernel(const float* xyz, long long n, float* out) { eval_stream_f32_body<true>(SpecEval(), xyz, n, out); }
#elif SDF_JIT_KERNEL == 7
extern "C" __global__ void __launch_bounds__(256)
jit_kernel(const double* xyz, long long n, double* out) { eval_stream_f64_body(SpecEval(), xyz, n, out); }
#else
#error "unknown SDF_JIT_KERNEL"
#endif
