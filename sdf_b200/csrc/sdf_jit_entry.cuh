// sdf_jit_entry.cuh -- kernel entry points of the run-time compiled translation unit.
// Included AFTER the generated `struct SpecEval` (sdfb200.cu: jit_source); SDF_JIT_KERNEL selects
// the one kernel this compilation emits, always under the name `jit_kernel`.
#pragma once

#if SDF_JIT_KERNEL == 0
extern "C" __global__ void __launch_bounds__(128)
jit_kernel(const GridDesc g, int sparse, unsigned char* flags, unsigned int* active_list, Counters* ctr) {
    skip_prepass_body(SpecEval(), g, sparse, flags, active_list, ctr);
}
#elif SDF_JIT_KERNEL >= 1 && SDF_JIT_KERNEL <= 4
extern "C" __global__ void __launch_bounds__(FUSED_THREADS, 1)
jit_kernel(const GridDesc g, const float tau, const unsigned int* active_list, unsigned int* batch_tris,
           Counters* ctr, float* tri_out, const unsigned long long tri_capacity) {
    fused_tile_body<32, SDF_JIT_KERNEL - 1>(SpecEval(), g, tau, active_list, batch_tris, ctr, tri_out, tri_capacity);
}
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
