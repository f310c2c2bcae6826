// fdes_b200 -- C ABI (include/fdes_b200.h): the drop-in FDES() symbol of the reference
// (src/FDESExport.cu:59-178) and the session API over the engine.
#include "../../include/fdes_b200.h"
#include "engine.h"
#include "emd.h"
#include "qsc.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <vector>
#include <algorithm>
#include <string>
#include <chrono>
#include <thread>
#include <exception>

using namespace fdes;

std::vector<int> fdes_b200_gpu_list(const char* spec, int first);

// A session owns one engine per GPU it runs on.  With several GPUs the work of a run is sharded
// over the engines of this ONE process (one host thread per device while they compute):
//   SHARD_CONFIGS   frozen-phonon configurations j of every measurement k (src/crystalMaker.cu:332);
//                   the partial sums are reduced onto the first device (Engine::reduce_from)
//   SHARD_REPLICAS  every engine holds the whole problem: measurements k (tilt / defocus series,
//                   src/crystalMaker.cu:324) or STEM probe positions are dealt out, nothing to reduce
enum { SHARD_NONE = 0, SHARD_CONFIGS = 1, SHARD_REPLICAS = 2 };
struct fdes_b200_sim {
    std::vector<std::unique_ptr<Engine>> engs;
    Engine* eng = nullptr;   // engs[0]: the engine of single-GPU sessions, the reduction root otherwise
    int shard = SHARD_NONE;
    Params params;   // as read (before sub-slicing)
    Atoms atoms;
    int m3_orig = 1;
};

// f(r) for r = 0 .. n-1 on n host threads (r = 0 on the caller's); the first exception is rethrown
template <class F>
static void parallel_for(int n, F f)
{
    if (n <= 1) { if (n == 1) f(0); return; }
    std::vector<std::exception_ptr> err(n);
    std::vector<std::thread> th;
    th.reserve(n - 1);
    for (int r = 1; r < n; r++)
        th.emplace_back([&, r] { try { f(r); } catch (...) { err[r] = std::current_exception(); } });
    try { f(0); } catch (...) { err[0] = std::current_exception(); }
    for (auto& t : th) t.join();
    for (auto& e : err) if (e) std::rethrow_exception(e);
}

static thread_local std::string g_err;

#define API_TRY try {
#define API_CATCH(ret)                                                                           \
    }                                                                                            \
    catch (const std::exception& e) { g_err = e.what(); return ret; }                            \
    catch (...) { g_err = "unknown error"; return ret; }

extern "C" {

const char* fdes_b200_last_error(void) { return g_err.c_str(); }
int fdes_b200_version(void) { return 100; }

int fdes_b200_parse_cnf(const char* cnf_path, int* dims, float* scalars, float* per_k,
                        float* atoms6_out, int max_atoms)
{
    API_TRY
    if (!cnf_path) throw std::runtime_error("cnf_path is NULL");
    Params p;
    Atoms at;
    if (!read_input(cnf_path, p, &at, false)) throw std::runtime_error(std::string("cannot read ") + cnf_path);
    set_sub_slices(p, sub_slice_ratio(p.d3, p.subSlTh));
    if (dims) {
        dims[0] = p.n1; dims[1] = p.n2; dims[2] = p.n3; dims[3] = p.m1; dims[4] = p.m2; dims[5] = p.m3;
        dims[6] = at.size(); dims[7] = (int)list_of_elements(at.Z).size(); dims[8] = p.frPh; dims[9] = p.mode;
    }
    if (scalars) {
        scalars[0] = p.lambda; scalars[1] = p.sigma; scalars[2] = p.gamma; scalars[3] = p.d1;
        scalars[4] = p.d2; scalars[5] = p.d3; scalars[6] = p.E0; scalars[7] = p.imPot;
    }
    if (per_k)
        for (int k = 0; k < p.n3; k++) {
            per_k[5 * k + 0] = p.tiltspec[2 * k]; per_k[5 * k + 1] = p.tiltspec[2 * k + 1];
            per_k[5 * k + 2] = p.tiltbeam[2 * k]; per_k[5 * k + 3] = p.tiltbeam[2 * k + 1];
            per_k[5 * k + 4] = p.defoci[k];
        }
    if (atoms6_out)
        for (int i = 0; i < at.size() && i < max_atoms; i++) {
            float* a = atoms6_out + 6 * (size_t)i;
            a[0] = (float)at.Z[i]; a[1] = at.xyz[3 * i]; a[2] = at.xyz[3 * i + 1]; a[3] = at.xyz[3 * i + 2];
            a[4] = at.dwf[i]; a[5] = at.occ[i];
        }
    return at.size();
    API_CATCH(-1)
}

int fdes_b200_write_used_cnf(const char* input_path, const char* out_path)
{
    API_TRY
    if (!input_path || !out_path) throw std::runtime_error("path is NULL");
    Params p;
    Atoms at;
    if (!read_input(input_path, p, &at, false)) throw std::runtime_error(std::string("cannot read ") + input_path);
    if (!write_cnf(out_path, p, at, 0)) throw std::runtime_error(std::string("cannot write ") + out_path);
    return 0;
    API_CATCH(-1)
}

int fdes_b200_qsc_scan(const char* qsc_path, int* nxy, float* xy_host, int max_probes, float* det_mrad_host,
                       int max_det)
{
    API_TRY
    if (!qsc_path) throw std::runtime_error("qsc_path is NULL");
    QscScan sc;
    if (!read_qsc_scan(qsc_path, sc)) throw std::runtime_error(std::string("cannot read ") + qsc_path);
    const int np = sc.nx * sc.ny, nd = (int)sc.det_mrad.size() / 2;
    if (nxy) { nxy[0] = sc.nx; nxy[1] = sc.ny; }
    if (xy_host) memcpy(xy_host, sc.xy.data(), sizeof(float) * 2 * (size_t)std::min(np, max_probes));
    if (det_mrad_host) memcpy(det_mrad_host, sc.det_mrad.data(), sizeof(float) * 2 * (size_t)std::min(nd, max_det));
    return nd;
    API_CATCH(-1)
}

int fdes_b200_write_emd(const char* input_path, const char* emd_path, const float* image_host,
                        const float* potential_host, int pot_slices, const float* exitwave_host)
{
    API_TRY
    if (!input_path || !emd_path) throw std::runtime_error("path is NULL");
    Params p;
    Atoms at;
    if (!read_input(input_path, p, &at, false)) throw std::runtime_error(std::string("cannot read ") + input_path);
    if (!write_emd(emd_path, p, at, image_host, potential_host, pot_slices, exitwave_host))
        throw std::runtime_error(std::string("cannot write ") + emd_path);
    return 0;
    API_CATCH(-1)
}

static std::unique_ptr<fdes_b200_sim> read_session(const char* cnf_path, const float* atoms6, int numAtoms)
{
    auto sim = std::make_unique<fdes_b200_sim>();
    if (!cnf_path) throw std::runtime_error("cnf_path is NULL");
    if (!read_input(cnf_path, sim->params, &sim->atoms, atoms6 != nullptr))
        throw std::runtime_error(std::string("cannot read ") + cnf_path);
    if (atoms6) {
        if (numAtoms <= 0) throw std::runtime_error("numAtoms must be positive");
        atoms_from_array(atoms6, numAtoms, sim->atoms);
    }
    sim->params.nAt = sim->atoms.size();
    sim->m3_orig = sim->params.m3;
    return sim;
}

fdes_b200_sim* fdes_b200_open_cnf(const char* cnf_path, const float* atoms6, int numAtoms,
                                  int gpu_index, int batch, int rank, int world, int want_exitwave)
{
    API_TRY
    auto sim = read_session(cnf_path, atoms6, numAtoms);
    EngineOptions opt;
    opt.gpu_index = gpu_index; opt.batch = batch; opt.rank = rank; opt.world = world;
    opt.want_exitwave = want_exitwave != 0;
    sim->engs.push_back(std::make_unique<Engine>(sim->params, sim->atoms, opt));
    sim->eng = sim->engs[0].get();
    return sim.release();
    API_CATCH(nullptr)
}

fdes_b200_sim* fdes_b200_open_multi(const char* cnf_path, const float* atoms6, int numAtoms,
                                    const int* gpu_indices, int ngpus, int batch, int want_exitwave)
{
    API_TRY
    if (!gpu_indices || ngpus < 1) throw std::runtime_error("need at least one GPU index");
    if (ngpus > MAX_PEERS + 1) throw std::runtime_error("at most 16 GPUs per session");
    auto sim = read_session(cnf_path, atoms6, numAtoms);
    const Params& p = sim->params;
    const int count = p.frPh > 0 ? p.frPh : 1;
    int n = 1;
    if (ngpus > 1 && count >= 2) { sim->shard = SHARD_CONFIGS; n = std::min(ngpus, count); }
    else if (ngpus > 1 && (p.n3 >= 2 || p.mode == 2)) { sim->shard = SHARD_REPLICAS; n = ngpus; }
    sim->engs.resize(n);
    fdes_b200_sim* raw = sim.get();
    parallel_for(n, [&](int r) {
        EngineOptions opt;
        opt.gpu_index = gpu_indices[r]; opt.batch = batch;
        opt.rank = raw->shard == SHARD_CONFIGS ? r : 0;
        opt.world = raw->shard == SHARD_CONFIGS ? n : 1;
        opt.want_exitwave = want_exitwave != 0;
        raw->engs[r] = std::make_unique<Engine>(raw->params, raw->atoms, opt);
    });
    sim->eng = sim->engs[0].get();
    return sim.release();
    API_CATCH(nullptr)
}

int fdes_b200_num_gpus(const fdes_b200_sim* sim) { return sim ? (int)sim->engs.size() : -1; }

void fdes_b200_close(fdes_b200_sim* sim) { delete sim; }

int fdes_b200_get_dims(const fdes_b200_sim* sim, int* d)
{
    API_TRY
    const Params& p = sim->eng->params();
    d[0] = p.n1; d[1] = p.n2; d[2] = p.n3; d[3] = p.m1; d[4] = p.m2; d[5] = p.m3;
    d[6] = sim->atoms.size(); d[7] = sim->eng->num_species(); d[8] = sim->eng->configs_total();
    d[9] = sim->eng->batch();
    return 0;
    API_CATCH(-1)
}

int fdes_b200_get_scalars(const fdes_b200_sim* sim, float* s)
{
    API_TRY
    const Params& p = sim->eng->params();
    s[0] = p.lambda; s[1] = p.sigma; s[2] = p.gamma; s[3] = p.d1; s[4] = p.d2; s[5] = p.d3;
    s[6] = p.E0; s[7] = p.imPot;
    return 0;
    API_CATCH(-1)
}

int fdes_b200_set_accumulators(fdes_b200_sim* sim, float* intensity_dev, float* exitwave_dev)
{
    API_TRY
    sim->eng->set_accumulators(intensity_dev, reinterpret_cast<cpx*>(exitwave_dev));
    return 0;
    API_CATCH(-1)
}

int fdes_b200_run_k(fdes_b200_sim* sim, int k)
{
    API_TRY
    sim->eng->run_k(k);
    return 0;
    API_CATCH(-1)
}

int fdes_b200_finish_k(fdes_b200_sim* sim, int k, float* image_host, float* exitwave_host)
{
    API_TRY
    sim->eng->finish_k(k, image_host, exitwave_host);
    return 0;
    API_CATCH(-1)
}

int fdes_b200_simulate(fdes_b200_sim* sim, float* image_host, float* exitwave_host)
{
    API_TRY
    const Params& p = sim->eng->params();
    const size_t n12 = (size_t)p.n1 * p.n2, m12 = (size_t)p.m1 * p.m2;
    const int n = (int)sim->engs.size();
    auto img_k = [&](int k) { return image_host ? image_host + k * n12 : nullptr; };
    auto ew_k = [&](int k) { return exitwave_host ? exitwave_host + 2 * k * m12 : nullptr; };
    if (n > 1 && sim->shard == SHARD_CONFIGS) {
        // every device runs its configurations of measurement k; device 0 adds the partial sums
        // (peer reads over NVLink) and runs the detector tail, as the reference does once per k
        // after its j loop (src/crystalMaker.cu:332-372)
        std::vector<Engine*> others;
        for (int r = 1; r < n; r++) others.push_back(sim->engs[r].get());
        for (int k = 0; k < p.n3; k++) {
            parallel_for(n, [&](int r) { sim->engs[r]->run_k(k); });
            sim->eng->reduce_from(others);
            sim->eng->finish_k(k, img_k(k), ew_k(k));
        }
    } else if (n > 1 && sim->shard == SHARD_REPLICAS && p.n3 > 1) {
        parallel_for(n, [&](int r) {
            for (int k = r; k < p.n3; k += n) {
                sim->engs[r]->run_k(k);
                sim->engs[r]->finish_k(k, img_k(k), ew_k(k));
            }
        });
    } else {
        for (int k = 0; k < p.n3; k++) {
            sim->eng->run_k(k);
            sim->eng->finish_k(k, img_k(k), ew_k(k));
        }
    }
    return 0;
    API_CATCH(-1)
}

int fdes_b200_potential_slices_count(const fdes_b200_sim* sim) { return sim ? sim->m3_orig : -1; }

int fdes_b200_potential(fdes_b200_sim* sim, float* out_host)
{
    API_TRY
    sim->eng->potential_slices(out_host);
    return 0;
    API_CATCH(-1)
}

int fdes_b200_jitter_next(fdes_b200_sim* sim, int k, float* xyz_host)
{
    API_TRY
    sim->eng->next_jittered_coords(k, xyz_host);
    return 0;
    API_CATCH(-1)
}

int fdes_b200_bin_atoms(fdes_b200_sim* sim, const float* xyz_host, int* bins_host)
{
    API_TRY
    sim->eng->bin_tuples(xyz_host, bins_host);
    return 0;
    API_CATCH(-1)
}

int fdes_b200_phase_grating(fdes_b200_sim* sim, const float* xyz_host, int slice, float* V_host)
{
    API_TRY
    if (slice < 0 || slice >= sim->eng->params().m3) throw std::runtime_error("slice out of range");
    sim->eng->phase_grating(xyz_host, slice, V_host);
    return 0;
    API_CATCH(-1)
}

int fdes_b200_exit_wave(fdes_b200_sim* sim, const float* xyz_host, int k, float* psi_host)
{
    API_TRY
    sim->eng->exit_wave(xyz_host, k, psi_host);
    return 0;
    API_CATCH(-1)
}

double fdes_b200_bench_configs(fdes_b200_sim* sim, int k, int configs)
{
    API_TRY
    return sim->eng->bench_configs(k, configs);
    API_CATCH(-1.0)
}

double fdes_b200_stem_scan(fdes_b200_sim* sim, int k, int nprobes, const float* xy_host, int ndet,
                           const float* det_mrad_host, float* out_host)
{
    API_TRY
    const int n = (int)sim->engs.size();
    if (n > 1 && nprobes > 0 && ndet > 0) {
        std::vector<double> ms(n, 0.0);
        if (sim->shard == SHARD_CONFIGS) {
            // every device scans all positions with its share of the frozen-phonon configurations
            // (weights 1/count inside); the detector signals are summed here in device order
            std::vector<std::vector<float>> part(n);
            parallel_for(n, [&](int r) {
                part[r].resize((size_t)nprobes * ndet);
                sim->engs[r]->stem_scan(k, nprobes, xy_host, ndet, det_mrad_host, part[r].data(), &ms[r]);
            });
            for (size_t i = 0; i < (size_t)nprobes * ndet; i++) {
                float s = part[0][i];
                for (int r = 1; r < n; r++) s += part[r][i];
                out_host[i] = s;
            }
        } else {
            // replicas: contiguous ranges of the raster per device, gathered in place
            parallel_for(n, [&](int r) {
                const int i0 = (int)((long long)nprobes * r / n), i1 = (int)((long long)nprobes * (r + 1) / n);
                if (i1 > i0)
                    sim->engs[r]->stem_scan(k, i1 - i0, xy_host + 2 * (size_t)i0, ndet, det_mrad_host,
                                            out_host + (size_t)i0 * ndet, &ms[r]);
            });
        }
        return *std::max_element(ms.begin(), ms.end());
    }
    double ms = 0.0;
    sim->eng->stem_scan(k, nprobes, xy_host, ndet, det_mrad_host, out_host, &ms);
    return ms;
    API_CATCH(-1.0)
}

int fdes_b200_time_sweeps(fdes_b200_sim* sim, int k, int batch, int reps, float* ms6)
{
    API_TRY
    if (reps < 1) throw std::runtime_error("reps must be >= 1");
    sim->eng->time_sweeps(k, batch, reps, ms6);
    return 0;
    API_CATCH(-1)
}

int fdes_b200_get_counters(fdes_b200_sim* sim, long long* c, int reset)
{
    API_TRY
    const EngineTimings& t = sim->eng->timings();
    c[0] = t.slices_executed; c[1] = t.kernel_launches; c[2] = sim->eng->band_columns(); c[3] = 0;
    if (reset) sim->eng->reset_timings();
    return 0;
    API_CATCH(-1)
}

static void ck(cudaError_t e, const char* what)
{
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

int fdes_b200_fft2d(float* data_host, int N, int dir, int gpu_index)
{
    API_TRY
    if (!fft_size_supported(N)) throw std::runtime_error("unsupported FFT size");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("no CUDA device");
    ck(cudaSetDevice(gpu_index), "cudaSetDevice");
    const size_t NN = (size_t)N * N;
    cpx *d = nullptr, *tw = nullptr;
    ck(cudaMalloc(&d, NN * sizeof(cpx)), "cudaMalloc");
    const std::vector<cpx> h = make_twiddles(N);
    ck(cudaMalloc(&tw, h.size() * sizeof(cpx)), "cudaMalloc");
    ck(cudaMemcpy(tw, h.data(), h.size() * sizeof(cpx), cudaMemcpyHostToDevice), "cudaMemcpy");
    ck(cudaMemcpy(d, data_host, NN * sizeof(cpx), cudaMemcpyHostToDevice), "cudaMemcpy");
    SweepGeom g{N, N, N, tw};
    RowOpts ro;
    launch_rows_fft(g, d, d, dir, ROW_STORE, ro, 1, 0);
    launch_cols_fft(g, d, d, dir, COL_PLAIN, nullptr, 1.f, 1, 0);
    ck(cudaDeviceSynchronize(), "fft2d kernels");
    ck(cudaMemcpy(data_host, d, NN * sizeof(cpx), cudaMemcpyDeviceToHost), "cudaMemcpy");
    cudaFree(d); cudaFree(tw);
    return 0;
    API_CATCH(-1)
}

int fdes_b200_sort_records(unsigned int* keys, int* cols, float* w, int n, int nkeys, int* rowptr,
                           int gpu_index)
{
    API_TRY
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("no CUDA device");
    ck(cudaSetDevice(gpu_index), "cudaSetDevice");
    SortBuffers sb{};
    int* rp = nullptr;
    const size_t nn = (size_t)std::max(n, 1);
    ck(cudaMalloc(&sb.keys, nn * 4), "cudaMalloc"); ck(cudaMalloc(&sb.keys_tmp, nn * 4), "cudaMalloc");
    ck(cudaMalloc(&sb.cols, nn * 4), "cudaMalloc"); ck(cudaMalloc(&sb.cols_tmp, nn * 4), "cudaMalloc");
    ck(cudaMalloc(&sb.w, nn * 4), "cudaMalloc"); ck(cudaMalloc(&sb.w_tmp, nn * 4), "cudaMalloc");
    ck(cudaMalloc(&sb.hist, 256 * (size_t)sort_num_blocks(n) * 4), "cudaMalloc");
    ck(cudaMalloc(&rp, ((size_t)nkeys + 1) * 4), "cudaMalloc");
    ck(cudaMemcpy(sb.keys, keys, (size_t)n * 4, cudaMemcpyHostToDevice), "cudaMemcpy");
    ck(cudaMemcpy(sb.cols, cols, (size_t)n * 4, cudaMemcpyHostToDevice), "cudaMemcpy");
    ck(cudaMemcpy(sb.w, w, (size_t)n * 4, cudaMemcpyHostToDevice), "cudaMemcpy");
    int bits = 0;
    while ((1LL << bits) <= (long long)nkeys) bits++;
    launch_radix_sort(sb, n, bits, 1, 0);
    launch_row_pointers(sb.keys, n, rp, nkeys, 1, 0);
    ck(cudaDeviceSynchronize(), "sort kernels");
    ck(cudaMemcpy(keys, sb.keys, (size_t)n * 4, cudaMemcpyDeviceToHost), "cudaMemcpy");
    ck(cudaMemcpy(cols, sb.cols, (size_t)n * 4, cudaMemcpyDeviceToHost), "cudaMemcpy");
    ck(cudaMemcpy(w, sb.w, (size_t)n * 4, cudaMemcpyDeviceToHost), "cudaMemcpy");
    ck(cudaMemcpy(rowptr, rp, ((size_t)nkeys + 1) * 4, cudaMemcpyDeviceToHost), "cudaMemcpy");
    cudaFree(sb.keys); cudaFree(sb.keys_tmp); cudaFree(sb.cols); cudaFree(sb.cols_tmp);
    cudaFree(sb.w); cudaFree(sb.w_tmp); cudaFree(sb.hist); cudaFree(rp);
    return 0;
    API_CATCH(-1)
}

}  // extern "C"

// "<count>" -> first, first + 1, ... (modulo the device count); "<i>,<j>,..." -> that list;
// empty / NULL -> {first}
std::vector<int> fdes_b200_gpu_list(const char* spec, int first)
{
    std::vector<int> out;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) ndev = 0;
    if (!spec || !spec[0]) { out.push_back(first); return out; }
    if (strchr(spec, ',')) {
        const char* q = spec;
        while (*q) {
            char* end = nullptr;
            const long v = strtol(q, &end, 10);
            if (end == q) break;
            out.push_back((int)v);
            q = *end == ',' ? end + 1 : end;
        }
    } else {
        int n = atoi(spec);
        if (ndev > 0 && n > ndev) n = ndev;
        for (int i = 0; i < std::max(1, n); i++) out.push_back(ndev > 0 ? (first + i) % ndev : first + i);
    }
    if (out.empty()) out.push_back(first);
    return out;
}

extern "C" {

int fdes_b200_parse_gpu_list(const char* spec, int first, int* out, int max_out)
{
    const std::vector<int> v = fdes_b200_gpu_list(spec, first);
    for (int i = 0; i < (int)v.size() && i < max_out; i++) out[i] = v[i];
    return (int)v.size();
}

// ---------------------------------------------------------------------------------------------
// drop-in export (reference src/FDESExport.cu:59-178)
// ---------------------------------------------------------------------------------------------
void FDES(int gpu_Index, int print_Level, char* input_name, char* image_name, char* emd_save_name,
          float* atomsArray, int numAtoms, float* dstImage)
{
    fprintf(stderr, "\n  fdes_b200: B200-native forward multislice behind the FDES interface\n\n");
    fprintf(stderr, "   input_name %s  \n", input_name ? input_name : "(null)");
    if (!input_name || !(strstr(input_name, ".emd") || strstr(input_name, ".cnf") || strstr(input_name, ".qsc"))) {
        fprintf(stderr, " \n input file %s error   \n", input_name ? input_name : "(null)");
        exit(0);   // reference: src/FDESExport.cu:100-101
    }
    if (numAtoms <= 0) exit(0);   // readAtomsFromArray, src/paramStructure.cu:306-307
    if (print_Level < 0 || print_Level > 2) {
        fprintf(stderr, " \n printLevel error %s  \n", input_name);
        exit(EXIT_FAILURE);
    }
    using clk = std::chrono::steady_clock;
    const bool timing = getenv("FDES_B200_TIMING") != nullptr;
    const auto t0 = clk::now();
    // FDES_B200_GPUS = "<count>" (devices gpu_Index, gpu_Index + 1, ...) or "<i>,<j>,..." shards the
    // run over several GPUs of this process; unset: the reference's single device (gpu_Index)
    const std::vector<int> gpus = fdes_b200_gpu_list(getenv("FDES_B200_GPUS"), gpu_Index);
    fdes_b200_sim* sim = gpus.size() > 1
        ? fdes_b200_open_multi(input_name, atomsArray, numAtoms, gpus.data(), (int)gpus.size(), 0, print_Level > 1)
        : fdes_b200_open_cnf(input_name, atomsArray, numAtoms, gpus.empty() ? gpu_Index : gpus[0], 0, 0, 1, print_Level > 1);
    if (!sim) {
        fprintf(stderr, " \n fdes_b200: %s \n", fdes_b200_last_error());
        exit(EXIT_FAILURE);
    }
    const auto t1 = clk::now();
    // side-effect file of getParams (src/paramStructure.cu:629-631), written by a helper thread
    // while the GPU works (formatting tens of thousands of atom lines takes milliseconds)
    // (readQsc writes ParamsUsedQsc.txt instead, src/rwQsc.cu:1084)
    // (readQsc writes ParamsUsedQsc.txt, readHdf5 ParamsUsedEmd.txt instead: src/rwQsc.cu:1084,
    // src/rwHdf5.cu:2565); .cnf and .qsc inputs also leave "config.emd" behind (src/FDESExport.cu:123, 141)
    const bool from_emd = strstr(input_name, ".emd") != nullptr;
    const char* used_name = from_emd ? "ParamsUsedEmd.txt"
                          : (is_qsc_name(input_name) && !strstr(input_name, ".cnf")) ? "ParamsUsedQsc.txt" : "dataFDES_used.cnf";
    std::thread cnf_writer([sim, gpu_Index, used_name, from_emd] {
        write_cnf(used_name, sim->params, sim->atoms, gpu_Index);
        if (!from_emd) write_emd("config.emd", sim->params, sim->atoms, nullptr, nullptr, 0, nullptr);
    });
    fprintf(stderr, "  Number of atoms %d \n", sim->atoms.size());
    const Params& p = sim->eng->params();
    const size_t n123 = (size_t)p.n1 * p.n2 * p.n3, m12 = (size_t)p.m1 * p.m2;
    // the caller's buffer receives the images directly (exportFormedimage, src/FDESExport.cu:162)
    std::vector<float> image_own, ew, pot;
    float* image = dstImage;
    if (!image) { image_own.resize(n123); image = image_own.data(); }
    if (print_Level > 1) ew.resize(2 * m12 * p.n3);
    auto fail = [&](void) {
        fprintf(stderr, " \n fdes_b200: %s \n", fdes_b200_last_error());
        cnf_writer.join();
        exit(EXIT_FAILURE);
    };
    if (fdes_b200_simulate(sim, image, ew.empty() ? nullptr : ew.data()) != 0) fail();
    const auto t2 = clk::now();
    if (print_Level > 0) {
        pot.resize(2 * m12 * (size_t)sim->m3_orig);
        if (fdes_b200_potential(sim, pot.data()) != 0) fail();
    }
    {
        // results file of buildMeasurements (src/crystalMaker.cu:402 -> writeHdf5, src/rwHdf5.cu:27-1084):
        // HDF5 bytes written by our own serialiser (emd.cpp; no libhdf5 in this build), on a helper
        // thread while this one writes the raw image file
        std::thread emd_writer;
        if (emd_save_name && emd_save_name[0])
            emd_writer = std::thread([&] {
                write_emd(emd_save_name, sim->params, sim->atoms, image, pot.empty() ? nullptr : pot.data(),
                          sim->m3_orig, ew.empty() ? nullptr : ew.data());
            });
        if (image_name && image_name[0]) write_binary(image_name, image, n123);
        if (emd_writer.joinable()) emd_writer.join();
    }
    const auto t3 = clk::now();
    cnf_writer.join();
    fdes_b200_close(sim);
    const auto t4 = clk::now();
    if (timing) {
        auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "  fdes_b200 timing [ms]: open %.2f  simulate %.2f  outputs %.2f  close %.2f  total %.2f\n",
                ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t0, t4));
    }
}

void fdes_b200_release_cache(void) { release_device_cache(); }

}  // extern "C"
