// launcher.cu -- program entry for Juzhen programs on the B200 backend (replaces cpp/launcher.cu:44-101).
//
// A Juzhen program defines `int compute()` (cpp/juzhen.hpp:8) and links this file for main().
// The reference's main() creates the memory pools and the global cuBLAS handle, reads NVIDIA_TF32,
// runs compute(), then tears everything down.  Here the device side is one call: jz_init() selects
// the device, sizes the launch geometry and reads the GEMM mode (NVIDIA_TF32=1 keeps its meaning:
// single-pass TF32 instead of the default fp32-accurate 3xTF32; JZ_GEMM_MODE=3xtf32|tf32|fp32).
// No cuBLAS handle is created unless the program was built with -DJZ_LEGACY_CUBLAS_HANDLE for code
// that calls cuBLAS itself (TransformerLayer, ml/layer.hpp:2896-2926); this backend never uses it.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "../cpp/juzhen.hpp"

static void describe_devices() {
    int sms = 0, major = 0, minor = 0;
    size_t mem = 0;
    if (jz_device_info(&sms, &major, &minor, &mem) != JZ_OK) {
        std::fprintf(stderr, "juzhen-b200: %s\n", jz_last_error());
        std::exit(1);
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    const char* modes[] = {"3xTF32 (fp32 accuracy)", "TF32", "fp32 SIMT"};
    std::cout << "GPU " << dev << ": " << prop.name << ", sm_" << major << minor << ", " << sms << " SMs, "
              << (mem >> 30) << " GiB" << std::endl;
    std::cout << "GEMM mode: " << modes[jz_get_gemm_mode() & 3] << std::endl;
}

int main() {
    std::cout << "Juzhen on juzhen-b200 (sm_100a backend, ABI v" << jz_abi_version() << ")" << std::endl;
    std::cout << "______________________________________________" << std::endl;
    if (jz_init(-1) != JZ_OK) {
        std::fprintf(stderr, "juzhen-b200: %s\n", jz_last_error());
        return 1;
    }
    describe_devices();
    std::cout << "______________________________________________" << std::endl << std::endl;

    int ret = 1;
    {
        // host pools first, device pool last: destroyed in reverse order when compute() has returned
        Memory<int> host_int;
        Memory<float> host_f32;
        Memory<double> host_f64;
#ifdef JZ_LEGACY_CUBLAS_HANDLE
        CuBLASErrorCheck(cublasCreate(&Matrix<CUDAfloat>::global_handle));
#endif
        Memory<CUDAfloat> device_pool;
        const auto t0 = std::chrono::steady_clock::now();
        { ret = compute(); }
        std::cout << std::endl;
        cudaDeviceSynchronize();
        if (const char* st = std::getenv("JZ_STATS"); st && *st && *st != '0') {
            // what the reference's static Profilers cannot show: device-side launch and pool behaviour
            size_t live = 0, cached = 0, mallocs = 0, hits = 0;
            jz_pool_stats(&live, &cached, &mallocs, &hits);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            std::cout << "jz_stats: compute() " << ms << " ms, " << jz_launch_count() << " kernel launches, pool: "
                      << mallocs << " device allocations, " << hits << " reuses, " << (live >> 20) << " MiB live, "
                      << (cached >> 20) << " MiB cached" << std::endl;
        }
#ifdef JZ_LEGACY_CUBLAS_HANDLE
        CuBLASErrorCheck(cublasDestroy(Matrix<CUDAfloat>::global_handle));
#endif
    }
    return ret;
}
