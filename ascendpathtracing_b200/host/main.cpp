// Drop-in host for the B200 build: the flow of the reference's src/main.cpp (device branch, :46-92) --
// size the buffers, read input/rays.bin + input/spheres.bin, copy to the device, launch render_do on a stream,
// synchronise, copy back, write output/color.bin -- on the CUDA runtime instead of ACL.  File names, sizes and
// the kernel-entry signature are the reference's; W/H/SAMPLES, compile-time in the reference (src/common.h:4-6),
// come from the command line / environment here.
//
//   render_gpu [--width W] [--height H] [--samples S] [--depth D] [--gen [--seed N | --counter-rng N]] [--ppm]
//
//   --gen   generate input/rays.bin and input/spheres.bin on the device first (replaces scripts/gen_data.py;
//           default = replay of its NumPy MT19937 seed-0 stream, bit-identical files)
//   --ppm   also resolve on the device and write output/color.ppm (replaces scripts/data_visualization.py)
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "data_utils.h"
#include "pt_arena.hpp"

static int arg_int(int argc, char **argv, const char *name, const char *env, int dflt) {
    for (int i = 1; i + 1 < argc; i++)
        if (std::strcmp(argv[i], name) == 0)
            return std::atoi(argv[i + 1]);
    if (const char *e = std::getenv(env))
        return std::atoi(e);
    return dflt;
}
static bool arg_flag(int argc, char **argv, const char *name) {
    for (int i = 1; i < argc; i++)
        if (std::strcmp(argv[i], name) == 0)
            return true;
    return false;
}

int main(int argc, char **argv) {
    PtParams cfg;
    ptb200_default_params(&cfg);
    cfg.width = arg_int(argc, argv, "--width", "PT_WIDTH", cfg.width);
    cfg.height = arg_int(argc, argv, "--height", "PT_HEIGHT", cfg.height);
    cfg.samples = arg_int(argc, argv, "--samples", "PT_SAMPLES", cfg.samples);
    cfg.depth = arg_int(argc, argv, "--depth", "PT_DEPTH", cfg.depth);
    const bool gen = arg_flag(argc, argv, "--gen"), ppm = arg_flag(argc, argv, "--ppm");
    const int mt_seed = arg_int(argc, argv, "--seed", "PT_SEED", 0);
    const int counter_seed = arg_int(argc, argv, "--counter-rng", "PT_COUNTER_RNG", -1);

    if (ptb200_device_count() < 1) {
        ERROR_LOG("no CUDA device: the B200 build has no CPU fallback (use run.sh -r cpu for the reference's cpu mode)");
        return 1;
    }

    uint32_t blockDim = 8;
    size_t elementNums = static_cast<size_t>(cfg.width) * cfg.height * 4 * cfg.samples;
    size_t inputRayByteSize = elementNums * sizeof(uint32_t) * 6;
    size_t inputSphereByteSize = 512;
    size_t outputColorByteSize = elementNums * sizeof(uint32_t) * 3;

    int32_t deviceId = 0;
    CHECK_CUDA(cudaSetDevice(deviceId));
    cudaStream_t stream = nullptr;
    CHECK_CUDA(cudaStreamCreate(&stream));

    uint8_t *rayHost, *sphereHost, *colorHost;
    CHECK_CUDA(cudaMallocHost(reinterpret_cast<void **>(&rayHost), inputRayByteSize));
    CHECK_CUDA(cudaMallocHost(reinterpret_cast<void **>(&sphereHost), inputSphereByteSize));
    CHECK_CUDA(cudaMallocHost(reinterpret_cast<void **>(&colorHost), outputColorByteSize));

    int rc = 0;
    try {
        const size_t imageBytes = static_cast<size_t>(cfg.width) * cfg.height * 3;
        ptb200::DeviceArena arena(inputRayByteSize + inputSphereByteSize + outputColorByteSize + imageBytes + elementNums * 16 + (1 << 20));
        auto rayDevice = arena.Alloc(inputRayByteSize);
        auto sphereDevice = arena.Alloc(inputSphereByteSize);
        auto colorDevice = arena.Alloc(outputColorByteSize);

        if (gen) {  // scripts/gen_data.py on the device
            CHECK_PTB200(ptb200_default_scene(reinterpret_cast<float *>(sphereHost)));
            if (counter_seed >= 0) {
                CHECK_PTB200(ptb200_gen_rays(&cfg, stream, nullptr, static_cast<uint64_t>(counter_seed), 0, cfg.width, rayDevice.Get<float>()));
            } else {
                std::vector<double> u(2 * elementNums);
                CHECK_PTB200(ptb200_mt19937_uniforms(static_cast<uint32_t>(mt_seed), 0, u.size(), u.data()));
                auto uDevice = arena.Alloc(u.size() * sizeof(double));
                CHECK_CUDA(cudaMemcpyAsync(uDevice.Get(), u.data(), u.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
                CHECK_PTB200(ptb200_gen_rays(&cfg, stream, uDevice.Get<double>(), 0, 0, cfg.width, rayDevice.Get<float>()));
                CHECK_CUDA(cudaStreamSynchronize(stream));
            }
            CHECK_CUDA(cudaMemcpyAsync(rayHost, rayDevice.Get(), inputRayByteSize, cudaMemcpyDeviceToHost, stream));
            CHECK_CUDA(cudaStreamSynchronize(stream));
            if (!WriteFile("./input/rays.bin", rayHost, inputRayByteSize) || !WriteFile("./input/spheres.bin", sphereHost, inputSphereByteSize))
                rc = 1;
        }

        size_t got = 0;
        if (!ReadFile("./input/rays.bin", got, rayHost, inputRayByteSize) || got != inputRayByteSize) {
            ERROR_LOG("input/rays.bin: expected %zu bytes for %dx%dx%d", inputRayByteSize, cfg.width, cfg.height, cfg.samples);
            rc = 1;
        }
        CHECK_CUDA(cudaMemcpyAsync(rayDevice.Get(), rayHost, inputRayByteSize, cudaMemcpyHostToDevice, stream));
        if (!ReadFile("./input/spheres.bin", got, sphereHost, inputSphereByteSize))
            rc = 1;
        CHECK_CUDA(cudaMemcpyAsync(sphereDevice.Get(), sphereHost, inputSphereByteSize, cudaMemcpyHostToDevice, stream));

        if (rc == 0) {
            CHECK_PTB200(ptb200_set_legacy_config(&cfg));  // the reference bakes these into the kernel (src/render.cpp:256)
            PtParams active;
            ptb200_get_legacy_config(&active);
            if (active.width == cfg.width && active.height == cfg.height && active.samples == cfg.samples && active.depth == cfg.depth)
                render_do(blockDim, nullptr, stream, rayDevice.Get(), sphereDevice.Get(), colorDevice.Get());
            else  // sizes outside the reference's tiling rule: the run-time entry has no such rule
                CHECK_PTB200(render_do_ex(&cfg, stream, rayDevice.Get(), sphereDevice.Get(), colorDevice.Get(), 0, -1));
            CHECK_CUDA(cudaStreamSynchronize(stream));

            CHECK_CUDA(cudaMemcpy(colorHost, colorDevice.Get(), outputColorByteSize, cudaMemcpyDeviceToHost));
            if (!WriteFile("./output/color.bin", colorHost, outputColorByteSize))
                rc = 1;

            if (ppm) {  // scripts/data_visualization.py on the device
                auto imageDevice = arena.Alloc(imageBytes);
                std::vector<uint8_t> image(imageBytes);
                CHECK_PTB200(ptb200_resolve(&cfg, stream, colorDevice.Get<float>(), 0, cfg.width, imageDevice.Get()));
                CHECK_CUDA(cudaMemcpyAsync(image.data(), imageDevice.Get(), imageBytes, cudaMemcpyDeviceToHost, stream));
                CHECK_CUDA(cudaStreamSynchronize(stream));
                if (ptb200_write_ppm("./output/color.ppm", cfg.width, cfg.height, image.data()) != PTB200_OK) {
                    ERROR_LOG("%s", ptb200_last_error());
                    rc = 1;
                } else {
                    INFO_LOG("Generate Result Image");
                }
            }
        }
    } catch (const std::exception &e) {
        ERROR_LOG("%s", e.what());
        rc = 1;
    }

    CHECK_CUDA(cudaFreeHost(rayHost));
    CHECK_CUDA(cudaFreeHost(sphereHost));
    CHECK_CUDA(cudaFreeHost(colorHost));
    CHECK_CUDA(cudaStreamDestroy(stream));
    if (cudaGetLastError() != cudaSuccess)
        rc = 1;
    return rc;
}
