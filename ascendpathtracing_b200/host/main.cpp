// Drop-in host for the B200 build: the flow of the reference's src/main.cpp (device branch, :46-92) --
// size the buffers, read input/rays.bin + input/spheres.bin, copy to the device, launch render_do on a stream,
// synchronise, copy back, write output/color.bin -- on the CUDA runtime instead of ACL.  File names, sizes and
// the kernel-entry signature are the reference's; W/H/SAMPLES, compile-time in the reference (src/common.h:4-6),
// come from the command line / environment here.
//
//   render_gpu [--width W] [--height H] [--samples S] [--depth D]
//              [--gen [--seed N | --counter-rng N] [--scene-kind default|smallpt|random:N[:SEED]]]
//              [--ppm] [--p6] [--materials [--max-depth N] [--rr-start N] [--mat-seed N] [--eps X]] [--bvh] [--gamma]
//              [--image] [--gpus N | --devices a,b,c] [--reps R] [--json FILE]
//
//   --gen         write input/spheres.bin (and, unless --image, input/rays.bin) on the device first: replaces
//                 scripts/gen_data.py; default ray stream = replay of its NumPy MT19937 seed-0 stream, bit-identical files
//   --scene-kind  which scene --gen writes: the reference's 8 spheres (512 bytes, the reference's own file), smallpt's 9
//                 spheres with materials, or BASELINE config C4's walls + light + N random spheres (11-row SoA, 44 bytes
//                 per column; include/ptb200.h: ptb200_scene_layout)
//   --ppm         also resolve on the device and write output/color.ppm (replaces scripts/data_visualization.py); --p6 = binary P6
//   --materials   DIFF/SPEC/REFR + Russian roulette instead of the reference's mirror kernel; --bvh: GPU-built sphere BVH
//   --image       production path: scene file in, output/color.ppm out, rays and per-path colours never leave the GPU
//                 (the only way to render frames whose rays.bin would not fit anywhere: 4K x 1024 spp = 204 GB)
//   --gpus N      drop-in mode: the N-way split of the path array the reference does over its 8 AI cores
//                 (src/render.cpp:24-27), one GPU per slice; --image: strided columns + peer-to-peer gather on the first GPU
//   --reps R      repeat the compute part R times and report the best (the first pass pays context and module loading)
// One JSON line with paths, milliseconds, Mpaths/s and Grays/s goes to stdout (and to --json FILE).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <sys/stat.h>

#include "data_utils.h"
#include "pt_arena.hpp"

namespace {

const char *arg_str(int argc, char **argv, const char *name, const char *env, const char *dflt) {
    for (int i = 1; i + 1 < argc; i++)
        if (std::strcmp(argv[i], name) == 0)
            return argv[i + 1];
    if (env != nullptr)
        if (const char *e = std::getenv(env))
            return e;
    return dflt;
}
long long arg_ll(int argc, char **argv, const char *name, const char *env, long long dflt) {
    const char *v = arg_str(argc, argv, name, env, nullptr);
    return v ? std::atoll(v) : dflt;
}
bool arg_flag(int argc, char **argv, const char *name) {
    for (int i = 1; i < argc; i++)
        if (std::strcmp(argv[i], name) == 0)
            return true;
    return false;
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

size_t file_size(const char *path) {
    struct stat sb;
    return (stat(path, &sb) == 0 && S_ISREG(sb.st_mode)) ? static_cast<size_t>(sb.st_size) : 0;
}

// --scene-kind: fills `scene` (host) and its layout; returns false on a bad kind
bool make_scene(const std::string &kind, std::vector<float> &scene, int32_t &count, int32_t &stride) {
    if (kind == "default") {  // the reference's file, byte for byte (scripts/gen_data.py:92-132)
        scene.assign(128, 0.0f);
        count = stride = 8;
        return ptb200_default_scene(scene.data()) == PTB200_OK;
    }
    if (kind == "smallpt") {
        scene.assign(176, 0.0f);
        count = 9, stride = 16;
        return ptb200_smallpt_scene(scene.data()) == PTB200_OK;
    }
    if (kind.rfind("random:", 0) == 0) {
        const char *s = kind.c_str() + 7;
        char *end = nullptr;
        const long n = std::strtol(s, &end, 10);
        const unsigned long seed = (end && *end == ':') ? std::strtoul(end + 1, nullptr, 10) : 12345ul;  // MT19937 seed 12345
        if (n < 0 || n > (1 << 24))
            return false;
        count = stride = static_cast<int32_t>(7 + n);
        scene.assign(static_cast<size_t>(11) * stride, 0.0f);
        return ptb200_random_scene(static_cast<int32_t>(n), static_cast<uint32_t>(seed), stride, scene.data()) == PTB200_OK;
    }
    return false;
}

bool write_p6(const char *path, int w, int h, const uint8_t *image) {
    FILE *f = fopen(path, "wb");
    if (f == nullptr)
        return false;
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    const size_t n = static_cast<size_t>(w) * h * 3;
    const bool ok = fwrite(image, 1, n, f) == n;
    return fclose(f) == 0 && ok;
}

}  // namespace

int main(int argc, char **argv) {
    PtParams cfg;
    ptb200_default_params(&cfg);
    cfg.width = static_cast<int32_t>(arg_ll(argc, argv, "--width", "PT_WIDTH", cfg.width));
    cfg.height = static_cast<int32_t>(arg_ll(argc, argv, "--height", "PT_HEIGHT", cfg.height));
    cfg.samples = static_cast<int32_t>(arg_ll(argc, argv, "--samples", "PT_SAMPLES", cfg.samples));
    cfg.depth = static_cast<int32_t>(arg_ll(argc, argv, "--depth", "PT_DEPTH", cfg.depth));
    const bool gen = arg_flag(argc, argv, "--gen"), ppm = arg_flag(argc, argv, "--ppm"), p6 = arg_flag(argc, argv, "--p6");
    const bool image_mode = arg_flag(argc, argv, "--image"), materials = arg_flag(argc, argv, "--materials");
    const bool use_bvh = arg_flag(argc, argv, "--bvh"), gamma = arg_flag(argc, argv, "--gamma");
    const int mt_seed = static_cast<int>(arg_ll(argc, argv, "--seed", "PT_SEED", 0));
    const long long counter_seed = arg_ll(argc, argv, "--counter-rng", "PT_COUNTER_RNG", -1);
    const std::string scene_kind = arg_str(argc, argv, "--scene-kind", "PT_SCENE_KIND", "default");
    const int reps = static_cast<int>(std::max(1LL, arg_ll(argc, argv, "--reps", "PT_REPS", 1)));
    const char *json_path = arg_str(argc, argv, "--json", nullptr, nullptr);
    PtMaterialParams mp;
    ptb200_default_material_params(&mp);
    mp.max_depth = static_cast<int32_t>(arg_ll(argc, argv, "--max-depth", "PT_MAX_DEPTH", mp.max_depth));
    mp.rr_start = static_cast<int32_t>(arg_ll(argc, argv, "--rr-start", "PT_RR_START", mp.rr_start));
    mp.seed = static_cast<uint64_t>(arg_ll(argc, argv, "--mat-seed", "PT_MAT_SEED", 0));
    if (const char *e = arg_str(argc, argv, "--eps", "PT_EPS", nullptr))
        mp.hit_epsilon = static_cast<float>(std::atof(e));

    // devices: --devices a,b,c or --gpus N (= 0 .. N-1)
    std::vector<int32_t> devices;
    if (const char *dl = arg_str(argc, argv, "--devices", "PT_DEVICES", nullptr)) {
        for (const char *s = dl; *s;) {
            devices.push_back(static_cast<int32_t>(std::strtol(s, const_cast<char **>(&s), 10)));
            if (*s == ',')
                s++;
            else if (*s)
                break;
        }
    } else {
        const int n = static_cast<int>(arg_ll(argc, argv, "--gpus", "PT_GPUS", 1));
        for (int i = 0; i < n; i++)
            devices.push_back(i);
    }
    const int n_dev = static_cast<int>(devices.size());

    if (ptb200_device_count() < 1) {
        ERROR_LOG("no CUDA device: the B200 build has no CPU fallback (use run.sh -r cpu for the reference's cpu mode)");
        return 1;
    }
    if (n_dev < 1 || n_dev > 16) {
        ERROR_LOG("%d GPUs requested: 1 to 16 supported", n_dev);
        return 1;
    }
    for (int32_t d : devices)  // (a device may be listed more than once with --devices: its shares then run one after the other)
        if (d < 0 || d >= ptb200_device_count()) {
            ERROR_LOG("GPU %d requested, %d visible", d, ptb200_device_count());
            return 1;
        }
    if (use_bvh && !materials) {
        ERROR_LOG("--bvh is the material kernel's scene representation: add --materials");
        return 1;
    }
    if (!image_mode && materials && n_dev > 1) {
        ERROR_LOG("--materials with --gpus > 1 needs --image (the drop-in ray-file mode splits only the reference kernel)");
        return 1;
    }

    uint32_t blockDim = 8;
    const size_t elementNums = static_cast<size_t>(cfg.width) * cfg.height * 4 * cfg.samples;
    const size_t inputRayByteSize = elementNums * sizeof(uint32_t) * 6;
    const size_t outputColorByteSize = elementNums * sizeof(uint32_t) * 3;
    const size_t imageBytes = static_cast<size_t>(cfg.width) * cfg.height * 3;

    int32_t deviceId = devices[0];
    CHECK_CUDA(cudaSetDevice(deviceId));
    cudaStream_t stream = nullptr;
    CHECK_CUDA(cudaStreamCreate(&stream));

    int rc = 0;
    // ---- the scene: written by --gen, then ALWAYS read back from input/spheres.bin like the reference's host (src/main.cpp:33)
    if (gen) {
        std::vector<float> scene;
        int32_t c = 0, s = 0;
        if (!make_scene(scene_kind, scene, c, s)) {
            ERROR_LOG("--scene-kind %s: expected default | smallpt | random:N[:SEED] (%s)", scene_kind.c_str(), ptb200_last_error());
            return 1;
        }
        if (!WriteFile("./input/spheres.bin", scene.data(), scene.size() * sizeof(float)))
            return 1;
    }
    const size_t inputSphereByteSize = file_size("./input/spheres.bin");  // 512 in the reference (src/main.cpp:24); size-derived here
    std::vector<float> sceneHostVec((inputSphereByteSize + 3) / 4 + 128, 0.0f);
    size_t got = 0;
    if (inputSphereByteSize == 0 || !ReadFile("./input/spheres.bin", got, sceneHostVec.data(), inputSphereByteSize)) {
        ERROR_LOG("input/spheres.bin missing or empty");
        return 1;
    }
    int32_t rows = 0;
    if (ptb200_scene_layout(sceneHostVec.data(), inputSphereByteSize, &cfg.sphere_count, &cfg.sphere_stride, &rows) != PTB200_OK) {
        ERROR_LOG("%s", ptb200_last_error());
        return 1;
    }
    if (cfg.sphere_count != 8 || inputSphereByteSize != 512)
        cfg.light_index = -1;  // the hard-coded light of the reference (rt_helper.h:776) only exists in its own scene
    if (inputSphereByteSize != 512 && !materials)
        WARN_LOG("a %d-sphere scene through the reference's mirror kernel has no light (index 7 is hard-coded there): consider --materials",
                 cfg.sphere_count);
    if (!use_bvh && cfg.sphere_count > 1024) {
        ERROR_LOG("%d spheres: the brute-force kernels take at most 1024, add --materials --bvh", cfg.sphere_count);
        return 1;
    }

    double best_ms = 0.0;
    uint64_t segments = 0;
    std::vector<double> dev_ms(1 + n_dev, 0.0);
    const char *mode = image_mode ? "image" : "dropin";

    if (image_mode) {
        // ---- production path: scene in, 8-bit frame out; nothing per-path ever leaves the GPUs ----
        uint8_t *imageHost = nullptr;
        CHECK_CUDA(cudaMallocHost(reinterpret_cast<void **>(&imageHost), imageBytes));
        const uint64_t seed = counter_seed >= 0 ? static_cast<uint64_t>(counter_seed) : 0;
        for (int rep = 0; rep < reps && rc == 0; rep++) {
            uint64_t stats[2] = {0, 0};
            const int prc = ptb200_render_image_multi(&cfg, materials ? &mp : nullptr, use_bvh ? 1 : 0, gamma ? 1 : 0, devices.data(), n_dev,
                                                      sceneHostVec.data(), seed, imageHost, stats, dev_ms.data());
            if (prc != PTB200_OK) {
                ERROR_LOG("%s", ptb200_last_error());
                rc = 1;
                break;
            }
            segments = stats[1];
            if (rep == 0 || dev_ms[0] < best_ms)
                best_ms = dev_ms[0];
        }
        if (rc == 0) {
            const bool ok = p6 ? write_p6("./output/color.ppm", cfg.width, cfg.height, imageHost)
                               : ptb200_write_ppm("./output/color.ppm", cfg.width, cfg.height, imageHost) == PTB200_OK;
            if (!ok) {
                ERROR_LOG("writing output/color.ppm failed: %s", ptb200_last_error());
                rc = 1;
            } else {
                INFO_LOG("Generate Result Image");
            }
        }
        CHECK_CUDA(cudaFreeHost(imageHost));
    } else {
        // ---- drop-in path: the reference's files (src/main.cpp:23-40) ----
        uint8_t *rayHost, *colorHost;
        CHECK_CUDA(cudaMallocHost(reinterpret_cast<void **>(&rayHost), inputRayByteSize));
        CHECK_CUDA(cudaMallocHost(reinterpret_cast<void **>(&colorHost), outputColorByteSize));
        try {
            const size_t sceneDevBytes = std::max<size_t>(512, sizeof(float) * 11 * static_cast<size_t>(cfg.sphere_stride));
            ptb200::DeviceArena arena(inputRayByteSize + sceneDevBytes + outputColorByteSize + imageBytes + (gen && counter_seed < 0 ? elementNums * 16 : 0) +
                                      (1 << 20));
            auto rayDevice = arena.Alloc(inputRayByteSize);
            auto sphereDevice = arena.Alloc(sceneDevBytes);
            auto colorDevice = arena.Alloc(outputColorByteSize);

            if (gen) {  // scripts/gen_data.py's rays on the device
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
                if (!WriteFile("./input/rays.bin", rayHost, inputRayByteSize))
                    rc = 1;
            }

            if (!ReadFile("./input/rays.bin", got, rayHost, inputRayByteSize) || got != inputRayByteSize) {
                ERROR_LOG("input/rays.bin: expected %zu bytes for %dx%dx%d", inputRayByteSize, cfg.width, cfg.height, cfg.samples);
                rc = 1;
            }

            for (int rep = 0; rep < reps && rc == 0; rep++) {
                const double t0 = now_ms();
                if (n_dev > 1) {
                    // the reference's blockDim-way split (src/render.cpp:24-27), one GPU per slice, host buffers in and out
                    if (ptb200_render_host_multi(&cfg, devices.data(), n_dev, reinterpret_cast<const float *>(rayHost), sceneHostVec.data(),
                                                 reinterpret_cast<float *>(colorHost), dev_ms.data()) != PTB200_OK) {
                        ERROR_LOG("%s", ptb200_last_error());
                        rc = 1;
                    }
                    segments = static_cast<uint64_t>(elementNums) * cfg.depth;
                } else {
                    CHECK_CUDA(cudaMemcpyAsync(rayDevice.Get(), rayHost, inputRayByteSize, cudaMemcpyHostToDevice, stream));
                    CHECK_CUDA(cudaMemcpyAsync(sphereDevice.Get(), sceneHostVec.data(), std::min(sceneDevBytes, sceneHostVec.size() * sizeof(float)),
                                               cudaMemcpyHostToDevice, stream));
                    if (materials) {
                        auto statDevice = arena.Alloc(256);
                        CHECK_CUDA(cudaMemsetAsync(statDevice.Get(), 0, 16, stream));
                        int mrc;
                        if (use_bvh) {
                            PtBvh *tree = nullptr;
                            mrc = ptb200_bvh_build(sphereDevice.Get(), cfg.sphere_count, cfg.sphere_stride, stream, &tree);
                            if (mrc == PTB200_OK)
                                mrc = render_do_mat_bvh(&cfg, &mp, tree, stream, rayDevice.Get(), colorDevice.Get(), 0, -1, 0, statDevice.Get<uint64_t>());
                            CHECK_CUDA(cudaStreamSynchronize(stream));
                            ptb200_bvh_destroy(tree);
                        } else {
                            mrc = render_do_mat(&cfg, &mp, stream, rayDevice.Get(), sphereDevice.Get(), colorDevice.Get(), 0, -1, 0, statDevice.Get<uint64_t>());
                        }
                        if (mrc != PTB200_OK) {
                            ERROR_LOG("%s", ptb200_last_error());
                            rc = 1;
                        }
                        CHECK_CUDA(cudaMemcpyAsync(&segments, statDevice.Get(), sizeof segments, cudaMemcpyDeviceToHost, stream));
                    } else {
                        // The reference bakes W/H/SAMPLES into the kernel (src/render.cpp:256) and asserts its tiling rule (:68-73);
                        // sizes that satisfy it go through the legacy entry, all others through the run-time sibling.
                        const bool legacy_ok = elementNums % 8 == 0 && (elementNums / 8) % 128 == 0 && ptb200_set_legacy_config(&cfg) == PTB200_OK;
                        if (legacy_ok)
                            render_do(blockDim, nullptr, stream, rayDevice.Get(), sphereDevice.Get(), colorDevice.Get());
                        else
                            CHECK_PTB200(render_do_ex(&cfg, stream, rayDevice.Get(), sphereDevice.Get(), colorDevice.Get(), 0, -1));
                        segments = static_cast<uint64_t>(elementNums) * cfg.depth;
                    }
                    CHECK_CUDA(cudaStreamSynchronize(stream));
                    CHECK_CUDA(cudaMemcpy(colorHost, colorDevice.Get(), outputColorByteSize, cudaMemcpyDeviceToHost));
                    dev_ms[1] = now_ms() - t0;
                }
                dev_ms[0] = now_ms() - t0;
                if (rep == 0 || dev_ms[0] < best_ms)
                    best_ms = dev_ms[0];
            }
            if (rc == 0 && !WriteFile("./output/color.bin", colorHost, outputColorByteSize))
                rc = 1;

            if (rc == 0 && ppm) {  // scripts/data_visualization.py on the device
                auto imageDevice = arena.Alloc(imageBytes);
                std::vector<uint8_t> image(imageBytes);
                if (n_dev > 1)  // the colours came back through host memory
                    CHECK_CUDA(cudaMemcpyAsync(colorDevice.Get(), colorHost, outputColorByteSize, cudaMemcpyHostToDevice, stream));
                CHECK_PTB200(ptb200_resolve(&cfg, stream, colorDevice.Get<float>(), 0, cfg.width, imageDevice.Get()));
                CHECK_CUDA(cudaMemcpyAsync(image.data(), imageDevice.Get(), imageBytes, cudaMemcpyDeviceToHost, stream));
                CHECK_CUDA(cudaStreamSynchronize(stream));
                const bool ok = p6 ? write_p6("./output/color.ppm", cfg.width, cfg.height, image.data())
                                   : ptb200_write_ppm("./output/color.ppm", cfg.width, cfg.height, image.data()) == PTB200_OK;
                if (!ok) {
                    ERROR_LOG("%s", ptb200_last_error());
                    rc = 1;
                } else {
                    INFO_LOG("Generate Result Image");
                }
            }
        } catch (const std::exception &e) {
            ERROR_LOG("%s", e.what());
            rc = 1;
        }
        CHECK_CUDA(cudaFreeHost(rayHost));
        CHECK_CUDA(cudaFreeHost(colorHost));
    }

    // ---- the report (SURVEY.md section 5 "metrics": keep the log macros and add one JSON timing / throughput line) ----
    if (rc == 0) {
        std::string j = "{\"ptb200\": \"render_gpu\", \"mode\": \"" + std::string(mode) + "\"";
        char buf[512];
        snprintf(buf, sizeof buf,
                 ", \"width\": %d, \"height\": %d, \"spp\": %d, \"depth\": %d, \"spheres\": %d, \"materials\": %s, \"bvh\": %s, \"gpus\": %d, \"reps\": %d"
                 ", \"paths\": %zu, \"segments\": %llu, \"ms\": %.3f, \"mpaths_per_s\": %.1f, \"grays_per_s\": %.3f, \"device_ms\": [",
                 cfg.width, cfg.height, 4 * cfg.samples, materials ? mp.max_depth : cfg.depth, cfg.sphere_count, materials ? "true" : "false",
                 use_bvh ? "true" : "false", n_dev, reps, elementNums, static_cast<unsigned long long>(segments), best_ms,
                 best_ms > 0 ? elementNums / best_ms / 1e3 : 0.0, best_ms > 0 ? segments / best_ms / 1e6 : 0.0);
        j += buf;
        for (int r = 0; r < n_dev; r++) {
            snprintf(buf, sizeof buf, "%s%.3f", r ? ", " : "", dev_ms[1 + r]);
            j += buf;
        }
        j += std::string("], \"timed\": \"") +
             (image_mode ? "scene upload, ray generation, trace, resolve, peer gather, image to host; best of reps"
                         : "H2D of rays and scene, trace, D2H of colours (pinned host buffers); best of reps") +
             "\"}";
        printf("%s\n", j.c_str());
        if (json_path != nullptr && !WriteFile(json_path, j.data(), j.size()))
            rc = 1;
    }

    CHECK_CUDA(cudaStreamDestroy(stream));
    if (cudaGetLastError() != cudaSuccess)
        rc = 1;
    return rc;
}
