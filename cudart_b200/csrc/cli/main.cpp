// cudart_b200/csrc/cli/main.cpp - `CudaRT`: the reference's command line on top of the B200 path.
//
//   ./CudaRT -f <scene> -s <spp> -w <width> -h <height> [--cpu] [--gpus N] [--progress N] [--dump-scene F] [--no-render]
//            [--metrics F]
//
// Same flags, defaults, `../.env` lookup with the `../meshes|settings|textures` fallback, same progress lines and
// the same three closing lines as the reference's main (main.cpp:24-188); output goes to <scene>.png in the
// working directory (`-o` is advertised by the reference's usage text but never parsed, D24 - same here).
// The B200Renderer below is the third implementation of the reference's Renderer seam (renderer.h:78-100):
// TerminalLauncher::render's call sequence (terminal_launcher.cpp:10-16) drives it unchanged.
// `--cpu` has no implementation in this repository: it hands over to an unmodified reference binary ($CUDART_REFERENCE,
// else baseline/_ref/CudaRT_ref next to this package) when one is installed and fails loudly otherwise.
#include <cuda_runtime.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/crt_abi.h"
#include "scene_io.h"

using namespace std;
using namespace std::chrono;
using namespace crtcli;

namespace {

[[noreturn]] void fatal(const string& what) {          // cuda_error_check.h:15-21 prints and exits
    cerr << what << endl;
    exit(1);
}
#define CRT(call) do { if ((call) != CRT_OK) fatal(string(#call) + ": " + crt_last_error()); } while (0)
#define CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) fatal(string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define NCCL(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) fatal(string(#call) + ": " + ncclGetErrorString(r_)); } while (0)

// The reference's Renderer contract (renderer.h:78-100) on top of the C ABI; one context per GPU, the image
// sharded as interleaved row tiles.  Read-out is the path's ONE exchange: every GPU resolves its own tiles to RGBA8
// (4 bytes per pixel instead of the accumulator's 16) and sends them to GPU 0 over NVLink (ncclSend / ncclRecv in one
// group, on the render streams, no host synchronisation in between); GPU 0 de-interleaves and copies to the host.
class B200Renderer {
public:
    bool useCuda = true;
    vector<uint8_t> h_img;                              // Renderer::h_imgPtr: W*H RGBA8

    B200Renderer(const SceneArrays& s, unsigned w, unsigned h, unsigned spp, int gpus) : width(w), height(h), samples(spp) {
        crt_scene_desc d{};
        d.tris = s.tris.data(); d.n_tris = (uint32_t)s.tris.size();
        d.nodes = s.nodes.data(); d.n_nodes = (uint32_t)s.nodes.size();
        d.mat_floats = s.mat_floats.data(); d.mat_indices = s.mat_indices.data(); d.n_mats = s.n_mats;
        d.light_indices = s.lights.data(); d.n_lights = (uint32_t)s.lights.size(); d.lights_area = s.lights_area;
        d.camera = &s.camera;
        d.tex_data = s.tex_data.data(); d.tex_dims = s.tex_dims.data(); d.tex_offsets = s.tex_offsets.data();
        d.n_tex = s.n_tex; d.tex_texels = (uint32_t)(s.tex_data.size() / 3);
        ctx.resize(gpus);
        streams.resize(gpus);
        for (int g = 0; g < gpus; g++) {
            CUDA(cudaSetDevice(g));
            CUDA(cudaStreamCreateWithFlags(&streams[g], cudaStreamNonBlocking));
            crt_options o{};
            o.struct_size = sizeof o;
            o.device = g; o.rank = (uint32_t)g; o.world = (uint32_t)gpus; o.stream = streams[g];
            CRT(crt_create(&d, w, h, &o, &ctx[g]));
        }
        if (gpus > 1) {
            comms.resize(gpus);
            vector<int> devs(gpus);
            for (int g = 0; g < gpus; g++) devs[g] = g;
            NCCL(ncclCommInitAll(comms.data(), gpus, devs.data()));
            void* p = nullptr;
            CRT(crt_resolve_local_rgba8_device(ctx[0], &p, &tile_bytes));
            CUDA(cudaSetDevice(0));
            CUDA(cudaMalloc(&gathered, tile_bytes * (size_t)gpus));
            // NCCL builds its channels lazily on the first exchange (hundreds of ms): do that here, with the other
            // one-time setup, so that the timed render (main.cpp:173-176 times Launcher::render only) does not pay for it
            exchange();
            for (int g = 0; g < gpus; g++) { CUDA(cudaSetDevice(g)); CUDA(cudaStreamSynchronize(streams[g])); }
        }
        h_img.assign((size_t)w * h * 4, 0);
    }
    ~B200Renderer() {
        for (size_t g = 0; g < ctx.size(); g++) {
            cudaSetDevice((int)g);
            if (g == 0 && gathered) cudaFree(gathered);
            crt_destroy(ctx[g]);
            if (g < comms.size()) ncclCommDestroy(comms[g]);
            cudaStreamDestroy(streams[g]);
        }
    }
    void renderOneSamplePerPixel(uint8_t*) { render(1); }
    void render(unsigned spp) {                          // batch form: `spp` calls of renderOneSamplePerPixel
        // One host thread per GPU: a render is thousands of launches per device and the driver's launch queue is
        // finite, so a single thread would feed the GPUs one after the other.
        vector<thread> workers;
        vector<string> errors(ctx.size());
        for (size_t g = 0; g < ctx.size(); g++)
            workers.emplace_back([&, g] {             // enqueue only: nothing waits for the device until the image is read
                if (crt_render(ctx[g], spp) != CRT_OK || crt_flush(ctx[g]) != CRT_OK) errors[g] = crt_last_error();
            });
        for (thread& t : workers) t.join();
        for (const string& e : errors) if (!e.empty()) fatal("crt_render: " + e);
        samplesRendered += spp;
    }
    void copyImageBytes(uint8_t*) {                      // Renderer::copyImageBytes: gather, resolve, gamma
        if (ctx.size() == 1) { CRT(crt_resolve_rgba8(ctx[0], h_img.data())); return; }
        exchange();
        CRT(crt_assemble_rgba8(ctx[0], gathered, h_img.data()));        // the one blocking call: de-interleave + D2H on GPU 0
        for (size_t g = 1; g < ctx.size(); g++) CRT(crt_sync(ctx[g]));
    }
    // every GPU: resolve its own tiles to RGBA8 on its render stream; GPUs 1..N-1 send them to GPU 0 (all enqueued, no host waits)
    void exchange() {
        vector<void*> local(ctx.size());
        for (size_t g = 0; g < ctx.size(); g++) {
            size_t bytes = 0;
            CRT(crt_resolve_local_rgba8_device(ctx[g], &local[g], &bytes));
            if (bytes != tile_bytes) fatal("copyImageBytes: ranks disagree about the tile buffer size");
        }
        CUDA(cudaSetDevice(0));
        CUDA(cudaMemcpyAsync(gathered, local[0], tile_bytes, cudaMemcpyDeviceToDevice, streams[0]));
        NCCL(ncclGroupStart());
        for (size_t g = 1; g < ctx.size(); g++) {
            NCCL(ncclRecv((char*)gathered + g * tile_bytes, tile_bytes, ncclUint8, (int)g, comms[0], streams[0]));
            NCCL(ncclSend(local[g], tile_bytes, ncclUint8, 0, comms[g], streams[g]));
        }
        NCCL(ncclGroupEnd());
    }
    uint8_t* getImgBytesPointer() { return h_img.data(); }
    void createMaterialsData() {}
    unsigned getWidth() const { return width; }
    unsigned getHeight() const { return height; }
    int getSamples() const { return (int)samples; }
    int getSamplesRendered() const { return (int)samplesRendered; }
    crt_counters counters() {
        crt_counters total{};
        for (crt_ctx* c : ctx) {
            crt_counters k{};
            CRT(crt_get_counters(c, &k));
            total.samples += k.samples; total.extend_rays += k.extend_rays; total.shadow_rays += k.shadow_rays;
            total.kernel_launches += k.kernel_launches; total.render_ms = max(total.render_ms, k.render_ms);
        }
        return total;
    }

private:
    unsigned width, height, samples;
    volatile unsigned samplesRendered = 0;
    vector<crt_ctx*> ctx;
    vector<cudaStream_t> streams;
    vector<ncclComm_t> comms;
    void* gathered = nullptr;                           // GPU 0: N x tile_bytes of RGBA8, rank-major
    size_t tile_bytes = 0;
};

// TerminalLauncher::render (terminal_launcher.cpp:10-16), batched: the renderer is told how many samples follow
// so that it can trace them in large wavefront batches instead of one launch sequence per sample.
void launch(B200Renderer& r, unsigned progress_every, const string& outFile, unsigned stop_after_writes) {
    uint8_t* img = r.getImgBytesPointer();
    if (progress_every == 0) {
        r.render((unsigned)r.getSamples());
    } else {
        // headless stand-in for the reference's --x window (window.cpp:17-68): the running image is written to the
        // output file every `progress_every` samples
        for (unsigned done = 0, total = (unsigned)r.getSamples(); done < total;) {
            const unsigned n = min(progress_every, total - done);
            r.render(n);
            done += n;
            if (done < total) {
                r.copyImageBytes(img);
                write_png_rgba8(outFile, r.h_img.data(), r.getWidth(), r.getHeight());
                cout << "\r" << done << " / " << total << " samples per pixel" << flush;
                if (stop_after_writes && --stop_after_writes == 0) {     // tests: what an interrupted run leaves behind
                    cout << endl;
                    exit(0);
                }
            }
        }
        cout << endl;
    }
    r.copyImageBytes(img);
}

string self_dir() {
    char buf[4096];
    const ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
    if (n <= 0) return ".";
    buf[n] = 0;
    string p(buf);
    return p.substr(0, p.find_last_of('/'));
}

}  // namespace

int main(int argc, char* argv[]) {
    int samples = 10, width = 640, height = 480, gpus = 1, progress = 0, stopAfterProgress = 0;
    string outFile = "out.png", sceneName = "cornell";
    string objDir, cameraDir, texturePath, dumpScene, metricsFile;
    const string envPath = "../.env";
    bool executeOnCpu = false, noRender = false;

    try {
        EnvFile env(envPath);
        objDir = env.meshes + "/";
        cameraDir = env.camera + "/";
        texturePath = env.textures + "/";
    } catch (const runtime_error&) {
        objDir = "../meshes/";
        cameraDir = "../settings/";
        texturePath = "../textures/";
    }
    string objPath = objDir + sceneName + ".obj", cameraPath = cameraDir + sceneName + "-camera.json";

    if (argc < 2) {
        cerr << "Usage: CudaRT <options>\n"
                "-o \t<output file>\n"
                "-s \t<number of samples>\tdefault:10\n"
                "-w \t<width>\tdefault:480px\n"
                "-h \t<height>\tdefault:320px\n"
                "-f \t<scene>\tdefault: cornell\n"
                "--cpu \t<flag to run sequential code on CPU only>\tdefault: false\n"
                "--gpus \t<number of B200s to shard the image over>\tdefault: 1\n"
                "--progress \t<write the running image every N samples>\tdefault: off\n";
        return 1;
    }
    vector<string> args(argv + 1, argv + argc);
    auto value_of = [&](const char* flag, const char* what, int& dst) {
        auto it = find(args.begin(), args.end(), flag);
        if (it < args.end() - 1) {
            try { dst = stoi(*(it + 1)); } catch (invalid_argument&) { cerr << "Invalid argument to " << what << "!" << endl; }
        }
    };
    value_of("-s", "-s", samples);
    value_of("-w", "-w", width);
    value_of("-h", "-h", height);
    value_of("--gpus", "--gpus", gpus);
    value_of("--progress", "--progress", progress);
    value_of("--stop-after-progress", "--stop-after-progress", stopAfterProgress);
    if (find(args.begin(), args.end(), "--cpu") < args.end()) executeOnCpu = true;
    if (find(args.begin(), args.end(), "--no-render") < args.end()) noRender = true;
    auto str_of = [&](const char* flag, string& dst) {
        auto it = find(args.begin(), args.end(), flag);
        if (it < args.end() - 1) dst = *(it + 1);
    };
    str_of("--dump-scene", dumpScene);
    str_of("--metrics", metricsFile);
    if (find(args.begin(), args.end(), "-f") < args.end() - 1) {
        sceneName = *(find(args.begin(), args.end(), "-f") + 1);
        objPath = objDir + sceneName + ".obj";
        cameraPath = cameraDir + sceneName + "-camera.json";
        outFile = sceneName + ".png";
    }

    if (executeOnCpu) {
        // The CPU path is the reference's own; this repository ships no CPU renderer and no fallback.
        const char* env_ref = getenv("CUDART_REFERENCE");
        const string ref = env_ref ? string(env_ref) : self_dir() + "/../baseline/_ref/CudaRT_ref";
        if (access(ref.c_str(), X_OK) != 0) fatal("--cpu: no reference binary at " + ref + " (set CUDART_REFERENCE to an unmodified CudaRT build); the B200 path has no CPU renderer and no fallback");
        execv(ref.c_str(), argv);
        fatal("--cpu: cannot exec " + ref);
    }

    cout << "Samples: " << samples << endl
         << "Width: " << width << endl
         << "Height: " << height << endl
         << "Obj path: " << objPath << endl
         << "Camera path: " << cameraPath << endl
         << "Output: " << outFile << endl;

    // CRT_CLI_TIMING=1: wall time of every phase on stderr (where the time to the image goes; not part of the reference's surface)
    const bool timing = getenv("CRT_CLI_TIMING") != nullptr;
    auto phase_start = high_resolution_clock::now();
    auto phase = [&](const char* name) {
        const auto now = high_resolution_clock::now();
        if (timing) cerr << "[timing] " << name << " " << duration_cast<microseconds>(now - phase_start).count() / 1e6 << " s" << endl;
        phase_start = now;
    };
    SceneArrays scene;
    try {
        cout << "Loading single .obj as scene from " << objPath << endl;
        ObjFile obj = load_obj(objPath);
        phase("load_obj");
        const crt::RefCamera cam = load_camera_json(cameraPath, (unsigned)width, (unsigned)height);
        scene = assemble_scene(obj, cam, texturePath);
        phase("assemble_scene (triangles, materials, textures, reference BVH, light list)");
        obj = ObjFile();
        phase("free the OBJ vertex soup");
    } catch (const exception& e) {
        fatal(string("terminate called after throwing an instance of 'std::runtime_error'\n  what():  ") + e.what());
    }
    for (size_t i = 0; i < scene.mesh_report.size(); i++)
        cout << "Mesh " << i << ": `" << scene.mesh_report[i].first << "`" << endl
             << "\t" << scene.mesh_report[i].second << " vertices | \t" << scene.mesh_report[i].second / 3 << " triangles" << endl;
    cout << "Total number of triangles:\t" << scene.tris.size() << endl;
    cout << "\nLaunching CudaRT version 0.1.0" << endl;
    if (!dumpScene.empty()) save_crts(dumpScene, scene, (unsigned)width, (unsigned)height);
    if (noRender) return 0;

    phase_start = high_resolution_clock::now();
    int devices = 0;
    cudaGetDeviceCount(&devices);
    phase("CUDA driver initialisation (cudaGetDeviceCount)");
    cout << "Got " << devices << " CUDA-capable devices" << endl;
    if (devices == 0) fatal("no CUDA device: the B200 path has no CPU fallback (use --cpu for the reference's own CPU renderer)");
    gpus = max(1, min(gpus, devices));

    phase_start = high_resolution_clock::now();
    unique_ptr<B200Renderer> renderer_owner(new B200Renderer(scene, (unsigned)width, (unsigned)height, (unsigned)samples, gpus));
    B200Renderer& renderer = *renderer_owner;
    phase("renderer (CUDA context, crt_create: derive + upload)");
    auto start = high_resolution_clock::now();
    launch(renderer, (unsigned)max(progress, 0), outFile, (unsigned)max(stopAfterProgress, 0));
    auto stop = high_resolution_clock::now();
    phase("render + read-back");
    write_png_rgba8(outFile, renderer.h_img.data(), (unsigned)width, (unsigned)height);     // Launcher::saveToImage
    phase("write PNG");

    const float elapsedTime = duration_cast<microseconds>(stop - start).count() / 1000000.0f;
    const int samplesRendered = renderer.getSamplesRendered();
    cout << "Rendered to " << outFile << " for " << samplesRendered << " samples per pixel. " << endl;
    cout << "Elapsed time in seconds = " << elapsedTime << endl;
    cout << "Average samples per pixel = " << samplesRendered / elapsedTime << endl;
    if (!metricsFile.empty()) {
        const crt_counters c = renderer.counters();
        ofstream m(metricsFile);
        m << "{\"scene\": \"" << sceneName << "\", \"width\": " << width << ", \"height\": " << height << ", \"spp\": " << samplesRendered
          << ", \"gpus\": " << gpus << ", \"seconds\": " << elapsedTime << ", \"msamples_per_s\": " << c.samples / elapsedTime / 1e6
          << ", \"mrays_per_s\": " << (c.extend_rays + c.shadow_rays) / elapsedTime / 1e6 << ", \"kernel_launches\": " << c.kernel_launches
          << ", \"device_render_ms\": " << c.render_ms << "}" << endl;
    }
    phase_start = high_resolution_clock::now();
    renderer_owner.reset();
    phase("teardown (crt_destroy: device buffers, streams, page-locked staging)");
    return 0;
}
