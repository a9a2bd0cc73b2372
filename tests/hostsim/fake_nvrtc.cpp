// Stand-in for libnvrtc (TEST INFRASTRUCTURE, tests/hostsim): "compiles" the translation unit sdfb200.cu's
// jit_source() generates with g++ against the SIMT stand-in and hands the resulting shared object back as the
// "cubin" (it is an ELF image, which is all the library's kernel cache checks).  The stand-in CUDA runtime's
// cudaLibraryLoadData dlopens it; a trampoline appended to the unit unpacks cudaLaunchKernel's argument array
// with the types of `jit_kernel`'s own signature.  So under `pytest --hostsim` the library's whole specialisation
// path runs on the host: source generation, candidate shapes, background compiles, the disk cache, argument
// marshalling, the share axis handed to the classification kernel.
//
//   $HOSTSIM_JIT_INCLUDE   directory of the patched device headers and the stand-in cuda_runtime.h the library
//                          was built with (tests/util_hostsim.py build_fake_nvrtc)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <vector>

namespace {
struct Program {
    std::string src, log;
    std::vector<char> image;
};
std::atomic<int> g_serial{0};

const char* kEpilogue =
    "\n// ---- appended by tests/hostsim/fake_nvrtc.cpp\n"
    "extern \"C\" void hostsim_jit_launch(unsigned grid, unsigned block, size_t smem, void** args, cudaStream_t stream) {\n"
    "    hostsim::launch_args(jit_kernel, grid, block, smem, args, stream);\n"
    "}\n";

std::string read_all(const std::string& path) {
    std::string out;
    if (FILE* f = fopen(path.c_str(), "rb")) {
        char buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
        fclose(f);
    }
    return out;
}
}  // namespace

extern "C" {

int nvrtcVersion(int* major, int* minor) { *major = 0; *minor = 1; return 0; }

int nvrtcCreateProgram(void** prog, const char* src, const char*, int, const char* const*, const char* const*) {
    Program* p = new Program();
    p->src = src;
    *prog = p;
    return 0;
}

int nvrtcDestroyProgram(void** prog) {
    delete (Program*)*prog;
    *prog = nullptr;
    return 0;
}

int nvrtcCompileProgram(void* prog, int, const char* const*) {
    Program* p = (Program*)prog;
    const char* inc = getenv("HOSTSIM_JIT_INCLUDE");
    if (!inc) { p->log = "fake nvrtc: HOSTSIM_JIT_INCLUDE not set"; return 1; }
    const char* tmp = getenv("TMPDIR") ? getenv("TMPDIR") : "/tmp";
    const std::string base = std::string(tmp) + "/hostsim_jit_" + std::to_string((long long)getpid()) + "_" +
                             std::to_string(g_serial.fetch_add(1));
    const std::string cpp = base + ".cpp", so = base + ".so", err = base + ".log";
    if (FILE* f = fopen(cpp.c_str(), "w")) { fputs(p->src.c_str(), f); fputs(kEpilogue, f); fclose(f); }
    else { p->log = "fake nvrtc: cannot write " + cpp; return 1; }
    // the unit shares hostsim::S with the library: it must be built like it (the sanitized build switches fibers differently)
    const char* san = getenv("HOSTSIM_SANITIZE") ? "-g -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer " : "";
    const std::string cmd = std::string("g++ -std=c++17 -O1 -fPIC -shared -ffp-contract=off -fno-strict-aliasing -Wno-unknown-pragmas ") + san +
                            "-Wl,-Bsymbolic-functions -I" + inc + " " + cpp + " -o " + so + " 2> " + err;
    const int rc = system(cmd.c_str());
    p->log = read_all(err);
    if (rc == 0) {
        const std::string bytes = read_all(so);
        p->image.assign(bytes.begin(), bytes.end());
    }
    remove(cpp.c_str()); remove(so.c_str()); remove(err.c_str());
    return rc == 0 && !p->image.empty() ? 0 : 1;
}

int nvrtcGetProgramLogSize(void* prog, size_t* n) { *n = ((Program*)prog)->log.size() + 1; return 0; }
int nvrtcGetProgramLog(void* prog, char* out) {
    const std::string& l = ((Program*)prog)->log;
    memcpy(out, l.c_str(), l.size() + 1);
    return 0;
}
int nvrtcGetCUBINSize(void* prog, size_t* n) { *n = ((Program*)prog)->image.size(); return 0; }
int nvrtcGetCUBIN(void* prog, char* out) {
    const std::vector<char>& im = ((Program*)prog)->image;
    memcpy(out, im.data(), im.size());
    return 0;
}

}  // extern "C"
