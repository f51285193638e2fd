// neural_color_transfer -- the reference's command line (NCT/main.cu:29-44, 546-590; NCT/CmdLine.cpp:21-57):
//   neural_color_transfer -m <model_dir> -i <input_root> -o <out_dir> -g <gpu_id> [-bds w] [-eps e] [-nl w] [-l w] [-w l]
// Flags may start with '-' or '/'; -h / -? / -help print the parameter list and exit with -1; an unknown flag prints
// "Unrecognized parameter" + the list and exits with -1.  Reads <model_dir>/vgg19/VGG_ILSVRC_19_layers.caffemodel and
// <input_root>/pairs.txt, writes <out_dir>/<content>_<style>_<bds>.png.
// Extension: -ngpu N processes the pair list on GPUs g .. g+N-1, one worker (own context) per GPU, pair i on worker
// i mod N; the reference is single-GPU (NCT/main.cu:563-565).  -inflight P keeps P pairs in flight per GPU (P contexts,
// streams and host threads per GPU; pair i on worker i mod (N * P)): the coarse pyramid levels and the solvers' small
// kernels do not fill a B200 on their own, so P = 4..6 raises the throughput of a long pair list by ~1.5x.  With P > 1
// the progress lines of different pairs interleave on stdout.  -resume 1 skips pairs whose output exists, -vis 1 writes the
// ENABLE_VIS debug images.  Exit code: 0 ok, -1 usage, 1 set-up failure, 2 at least one pair failed.  Images are PNG only.  -engine selects the convolution engine (the reference has
// whatever algorithm cuDNN picks): 3 = tensor cores, exact fixed point (tcgen05 kind::i8 digit planes, INT32 accumulation;
// default -- a pure function of the inputs, bit-identical to the oracle's fixed-point features); 2 = tensor cores, 3xTF32;
// 1 = plain TF32; 0 = FP32 CUDA cores in the canonical FP32 summation order (bit-identical to the oracle's FP32 features).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "../../../include/nct.h"

extern "C" int nct_vgg19_load_caffemodel(nct_ctx *ctx, const char *path);
extern "C" int nct_vgg19_set_engine(nct_ctx *ctx, int engine);

struct Param { const char *arg; const char *desc; int kind; void *dst; };  // kind 0 string, 1 int, 2 double

static bool is_arg(const char *s) { return s && (s[0] == '-' || s[0] == '/') && s[1] != 0; }

static void do_help(const char *prog, const std::vector<Param> &params)
{
    printf("Running: %s\n", prog);
    for (const Param &p : params) printf("    -%s %s\n", p.arg, p.desc);
}

int main(int argc, char **argv)
{
    nct_config cfg;
    nct_config_default(&cfg);
    std::string model_dir, input_dir, output_dir;
    int gpu_id = 0, ngpu = 1, inflight = 1, engine = 3, resume = 0, vis = 0;
    std::vector<Param> params = {
        {"m", "Directory of network models.", 0, &model_dir},
        {"i", "Input directory of content and style images (PNG) and pairs.txt.", 0, &input_dir},
        {"o", "Output directory of result images.", 0, &output_dir},
        {"g", "GPU ID (default: 0).", 1, &gpu_id},
        {"bds", "Weight of reverse color in BDS voting (default: 2.0).", 2, &cfg.bds_weight},
        {"eps", "Eps is used to avoid dividing zero (default: 0.6 with range in [0-255]).", 2, &cfg.var_eps},
        {"nl", "Weight of nonlocal constraint (default: 2.0).", 2, &cfg.nonlocal_weight},
        {"l", "Weight of local constraint (default: 0.125).", 2, &cfg.local_weight},
        {"w", "Initial value of WLS weight (default: 0.024).", 2, &cfg.wls_lambda_init},
        {"ngpu", "Number of GPUs to spread the pair list over, starting at -g (default: 1).", 1, &ngpu},
        {"inflight", "Pairs processed concurrently per GPU (default: 1).", 1, &inflight},
        {"resume", "1: skip pairs whose output file already exists (restart of an interrupted list; default: 0).", 1, &resume},
        {"vis", "1: also write the per-level debug images of the reference's ENABLE_VIS build into the output directory (default: 0).", 1, &vis},
        {"engine", "Convolution engine: 0 FP32 CUDA cores, 1 tcgen05 TF32, 2 tcgen05 3xTF32, 3 tcgen05 INT8 exact fixed point (default: 3; 0 and 3 are bit-exact against the oracle).", 1, &engine},
    };
    int i = 1;
    while (i < argc) {
        if (!is_arg(argv[i])) { ++i; continue; }  // stray file arguments are collected and ignored by the reference
        const std::string a(argv[i] + 1);
        if (a == "h" || a == "?" || a == "help") { do_help(argv[0], params); return -1; }
        bool processed = false;
        for (const Param &p : params) {
            if (a != p.arg) continue;
            if (i + 1 < argc) {
                if (p.kind == 0) *(std::string *)p.dst = argv[i + 1];
                else if (p.kind == 1) *(int *)p.dst = atoi(argv[i + 1]);
                else *(double *)p.dst = atof(argv[i + 1]);
                ++i;
            }
            ++i;
            processed = true;
            break;
        }
        if (!processed) {
            printf("Unrecognized parameter: %s\n\n", argv[i]);
            do_help(argv[0], params);
            return -1;
        }
    }
    if (ngpu < 1) ngpu = 1;
    if (inflight < 1) inflight = 1;
    const int nworkers = ngpu * inflight;
    const std::string weights = model_dir + "/vgg19/VGG_ILSVRC_19_layers.caffemodel";
    std::vector<nct_ctx *> ctxs((size_t)nworkers, nullptr);
    for (int r = 0; r < nworkers; ++r) {
        const int dev = gpu_id + r % ngpu;  // worker r -> GPU r mod N, so consecutive pairs go to different GPUs
        if (nct_create(dev, &ctxs[r]) != NCT_OK) {
            fprintf(stderr, "Error: cannot create a context on GPU %d (libnct needs an sm_100 device; there is no CPU path).\n", dev);
            return 1;
        }
        if (r == 0) printf("The number of device is: %d, set device %d.\n", ngpu, gpu_id);
        if (nct_vgg19_load_caffemodel(ctxs[r], weights.c_str()) != NCT_OK || nct_vgg19_set_engine(ctxs[r], engine) != NCT_OK) {
            fprintf(stderr, "Error: %s\n", nct_last_error(ctxs[r]));
            return 1;
        }
    }
    std::vector<std::thread> workers;
    std::vector<int> done((size_t)nworkers, 0), failed((size_t)nworkers, 0), skipped((size_t)nworkers, 0), rcs((size_t)nworkers, 0);
    const int flags = (resume ? NCT_RUN_RESUME : 0) | (vis ? NCT_RUN_VIS : 0);
    for (int r = 0; r < nworkers; ++r)
        workers.emplace_back([&, r]() {
            rcs[r] = nct_run_pairs_ex(ctxs[r], input_dir.c_str(), output_dir.c_str(), &cfg, r, nworkers, flags, &done[r], &failed[r], &skipped[r]);
        });
    for (auto &t : workers) t.join();
    int rc = 0, nfailed = 0;
    for (int r = 0; r < nworkers; ++r) {
        if (rcs[r]) { fprintf(stderr, "Error: %s\n", nct_last_error(ctxs[r])); rc = 1; }
        nfailed += failed[r];
        nct_destroy(ctxs[r]);
    }
    // the reference has no error path at all (a bad image is reported and skipped, the exit code stays 0); a batch tool
    // that silently drops pairs is worse: 2 = finished, but at least one pair failed
    if (!rc && nfailed) { fprintf(stderr, "Error: %d pair(s) failed.\n", nfailed); rc = 2; }
    return rc;
}
