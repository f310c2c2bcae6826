/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * Harness that links the UNMODIFIED reference objects (compiled in place from
 * /root/reference by oracle/Makefile into oracle/_ref/) and drives the
 * reference's own functions, so that
 *   (1) the numpy oracle (oracle/fdes_oracle.py) can be pinned against real
 *       reference output (golden vectors under tests/golden/), and
 *   (2) bench.py --impl reference can time the reference's stock code path
 *       (cuFFT + cuBLAS + 1-thread-per-pixel kernels) on the same B200.
 *
 * This file contains no reference source: it only *calls* reference entry
 * points with external linkage:
 *   getParams            src/paramStructure.cu:588
 *   readQsc              src/rwQsc.cu (include/rwQsc.h)
 *   buildMeasurements    src/crystalMaker.cu:227
 *   phaseGrating         src/crystalMaker.cu:507
 *   forwardPropagation   src/multisliceSimulation.cu:538
 *   incomingWave         src/multisliceSimulation.cu:563
 *   atomJitter           src/crystalMaker.cu:456
 *   tiltCoordinates      src/crystalMaker.cu:427
 *   listOfElements       src/crystalMaker.cu:539
 *   diffractionPattern   src/crystalMaker.cu:700
 *   applyLensFunction    src/multisliceSimulation.cu:614
 *   addNoiseAndMtf       src/crystalMaker.cu:579
 * and it defines the file-scope globals that src/FDES.cu:40-57 would define
 * (FDES.cu itself is not linked because it owns main()).
 *
 * Modes
 *   ref_harness run   <input.{cnf,qsc}> <outdir> <print_level>
 *       stock buildMeasurements(); the writeHdf5 stub dumps image /
 *       exit wave / potential as raw float32 into <outdir>.
 *   ref_harness trace <input.{cnf,qsc}> <outdir> <max_slices_to_dump>
 *       re-plays the driver loop of src/crystalMaker.cu:324-373 for k=0 with
 *       the reference's own functions and dumps V, psi per slice, the jittered
 *       coordinates per phonon configuration, I_d and J.
 *   ref_harness time  <input.{cnf,qsc}> <reps> <configs_per_rep>
 *       CUDA-event time of the slice loop (phaseGrating+forwardPropagation),
 *       prints one JSON line.
 *   ref_harness e2e   <input.cnf> <reps> <warmup> [time budget in seconds]
 *       whole-call wall time of the reference's exported entry point, replayed
 *       step by step as src/FDESExport.cu:59-178 does it (atoms handed over as a
 *       host array, image copied back into a host buffer): getParams ->
 *       readAtomsFromArray -> buildMeasurements -> image copy -> frees.  The
 *       exported FDES() itself is not called because it copies the image out of
 *       a buffer buildMeasurements has already freed (src/crystalMaker.cu:414 vs
 *       src/FDESExport.cu:162).  Prints one JSON line.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include <time.h>
#include <sys/stat.h>
#include <cuda_runtime.h>
#include <cufft.h>
#include <cublas_v2.h>
#include <curand_kernel.h>

#include "paramStructure.h"
#include "crystalMaker.h"
#include "multisliceSimulation.h"
#include "complexMath.h"
#include "rwQsc.h"
#include "rwHdf5.h"

/* globals normally owned by src/FDES.cu:40-57 */
const char* default_txt_name = "dataFDES.cnf";
const char* default_emd_name = "config.emd";
const char* default_qsc_name = "test.qsc";
const char* default_image_name = "Measurements.bin";
const char* default_emd_save_name = "results.emd";
int MATLAB_TILT_COMPATIBILITY = 0;
int NO_FUNCTION_EVALS_FDES = 0;
int NO_DERIVATIVE_EVALS_FDES = 0;
int printLevel = 0;
int confOption = -1;
int gpu_index = 0;
float version = 0.1f;
bool atomsFromExternal = false;
float* image = NULL;
float* potential = NULL;
float* exitwave = NULL;

static std::string g_outdir = ".";
static float* g_e2e_dst = NULL;   /* e2e mode: host buffer that receives the image */

static void dumpRaw(const std::string& name, const void* p, size_t bytes)
{
    std::string path = g_outdir + "/" + name;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "ref_harness: cannot write %s\n", path.c_str()); exit(3); }
    fwrite(p, 1, bytes, f);
    fclose(f);
}

static void dumpDev(const std::string& name, const void* d, size_t bytes)
{
    std::vector<char> h(bytes);
    cudaMemcpy(h.data(), d, bytes, cudaMemcpyDeviceToHost);
    dumpRaw(name, h.data(), bytes);
}

/* ---- stand-ins for src/rwHdf5.cu (libhdf5 is absent): signatures from
 *      include/rwHdf5.h:45-51.  The 9-argument overload is where the stock
 *      driver hands over its results (src/crystalMaker.cu:402). ---- */
void writeHdf5(const char*, params_t*, int**, float**, float**, float**) {}
void writeHdf5(const char*, float*, float*, params_t*) {}
void writeHdf5(const char*, float* img, float* pot, float* ew, params_t* params,
               int**, float**, float**, float**)
{
    const size_t n123 = (size_t)params->IM.n1 * params->IM.n2 * params->IM.n3;
    const size_t m12 = (size_t)params->IM.m1 * params->IM.m2;
    if (g_e2e_dst) { memcpy(g_e2e_dst, img, n123 * sizeof(float)); return; }
    dumpRaw("image.f32", img, n123 * sizeof(float));
    if (printLevel > 1 && ew)
        dumpRaw("exitwave.f32", ew, 2 * m12 * params->IM.n3 * sizeof(float));
    /* NOTE src/crystalMaker.cu:381-397: the potential buffer is only filled
     * when sub-slicing / phonons / tilt force the recomputation. */
    if (printLevel > 0 && pot)
        dumpRaw("potential.f32", pot, 2 * m12 * params->IM.m3 * sizeof(float));
    FILE* f = fopen((g_outdir + "/meta.txt").c_str(), "w");
    fprintf(f, "n1 %d\nn2 %d\nn3 %d\nm1 %d\nm2 %d\nm3 %d\nd1 %.9g\nd2 %.9g\nd3 %.9g\n"
               "lambda %.9g\nsigma %.9g\ngamma %.9g\nmode %d\nfrPh %d\nnAt %d\n",
            params->IM.n1, params->IM.n2, params->IM.n3, params->IM.m1, params->IM.m2,
            params->IM.m3, params->IM.d1, params->IM.d2, params->IM.d3, params->EM.lambda,
            params->EM.sigma, params->EM.gamma, params->IM.mode, params->IM.frPh,
            params->SAMPLE.nAt);
    fclose(f);
}
bool readHdf5(const char*, params_t**, int**, float**, float**, float**)
{
    fprintf(stderr, "ref_harness: .emd input unavailable (no libhdf5 in this image)\n");
    return false;
}

static std::string absPath(const char* p)
{
    if (p[0] == '/') return p;
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return p;
    return std::string(cwd) + "/" + p;
}

static bool loadInput(const std::string& in, params_t** params, int** Z_d, float** xyz_d,
                      float** DWF_d, float** occ_d)
{
    if (strstr(in.c_str(), ".qsc")) { confOption = 2; return readQsc(in.c_str(), params, Z_d, xyz_d, DWF_d, occ_d); }
    if (strstr(in.c_str(), ".cnf")) { confOption = 1; return getParams(in.c_str(), params, Z_d, xyz_d, DWF_d, occ_d); }
    fprintf(stderr, "ref_harness: unsupported input %s\n", in.c_str());
    return false;
}

/* state shared by trace/time: mirrors the set-up part of
 * src/crystalMaker.cu:227-322 by calling the reference's own helpers. */
struct Rig {
    params_t* params; params_t* params_d;
    int* Z_d; float *xyz_d, *DWF_d, *occ_d, *xyzTO_d, *xyzFP_d, *J_d;
    cufftComplex *V_d, *I_d, *psi, *t, *frProp;
    curandState *dwfState_d, *poissonState_d;
    cufftHandle planB; int Zlist[103]; int nZ, nAt, count, m12;
};

static void rigSetup(Rig& r)
{
    params_t* p = r.params;
    r.nAt = p->SAMPLE.nAt;
    setDeviceParams(&r.params_d, p);
    float ratio = subSliceRatio(p->IM.d3, p->IM.subSlTh);
    setSubSlices(p, r.params_d, ratio);
    r.m12 = p->IM.m1 * p->IM.m2;
    const size_t cb = (size_t)r.m12 * sizeof(cufftComplex);
    cudaMalloc(&r.xyzTO_d, r.nAt * 3 * sizeof(float));
    cudaMalloc(&r.xyzFP_d, r.nAt * 3 * sizeof(float));
    cudaMalloc(&r.V_d, cb); cudaMalloc(&r.I_d, cb); cudaMalloc(&r.psi, cb);
    cudaMalloc(&r.t, cb); cudaMalloc(&r.frProp, cb);
    cudaMalloc(&r.J_d, (size_t)p->IM.n1 * p->IM.n2 * p->IM.n3 * sizeof(float));
    cudaMalloc(&r.dwfState_d, 3 * (size_t)r.nAt * sizeof(curandState));
    cudaMalloc(&r.poissonState_d, (size_t)r.m12 * sizeof(curandState));
    cublasScopy(p->CU.cublasHandle, r.nAt * 3, r.xyz_d, 1, r.xyzTO_d, 1);
    tiltCoordinates(r.xyzTO_d, r.nAt, p->IM.specimen_tilt_offset_x, p->IM.specimen_tilt_offset_y,
                    p->IM.specimen_tilt_offset_z, p);
    r.nZ = listOfElements(r.Zlist, r.nAt, r.Z_d);
    setCufftPlanBatch(&r.planB, p);
    if (p->IM.frPh > 0)
        setupCurandState_d<<<myGSize(3 * r.nAt), myBSize(3 * r.nAt)>>>(r.dwfState_d, 1, 3 * r.nAt);
    if (p->IM.pD > FLT_MIN)
        setupCurandState_d<<<myGSize(r.m12), myBSize(r.m12)>>>(r.poissonState_d, 1 + p->IM.n3, r.m12);
    r.count = p->IM.frPh > 0 ? p->IM.frPh : 1;
    cudaDeviceSynchronize();
}

static int modeTrace(const std::string& in, int maxDump)
{
    Rig r; memset(&r, 0, sizeof r);
    if (!loadInput(in, &r.params, &r.Z_d, &r.xyz_d, &r.DWF_d, &r.occ_d)) return 2;
    rigSetup(r);
    params_t* p = r.params;
    const size_t cb = (size_t)r.m12 * sizeof(cufftComplex);
    cufftComplex alpha; alpha.x = 1.f / (float)r.count; alpha.y = 0.f;
    cufftComplex* ew_d; cudaMalloc(&ew_d, cb);
    const int k = 0;
    initialValues<<<p->CU.gS2D * 2, p->CU.bS>>>(r.I_d, r.m12, 0.f, 0.f);
    initialValues<<<p->CU.gS2D * 2, p->CU.bS>>>(ew_d, r.m12, 0.f, 0.f);
    cublasScopy(p->CU.cublasHandle, r.nAt * 3, r.xyzTO_d, 1, r.xyz_d, 1);
    tiltCoordinates(r.xyz_d, r.nAt, p->IM.tiltspec[2 * k], p->IM.tiltspec[2 * k + 1], 0.f, p);
    char name[256];
    for (int j = 0; j < r.count; j++) {
        incomingWave(r.psi, k, p, r.params_d);
        if (j == 0) dumpDev("psi_in.c64", r.psi, cb);
        cublasScopy(p->CU.cublasHandle, r.nAt * 3, r.xyz_d, 1, r.xyzFP_d, 1);
        if (p->IM.frPh > 0) atomJitter(r.xyzFP_d, r.dwfState_d, r.nAt, r.DWF_d);
        snprintf(name, sizeof name, "xyz_cfg%03d.f32", j);
        dumpDev(name, r.xyzFP_d, r.nAt * 3 * sizeof(float));
        for (int s = 0; s < p->IM.m3; s++) {
            phaseGrating(r.V_d, r.nAt, r.nZ, p, r.params_d, r.xyzFP_d, p->SAMPLE.imPot, r.Z_d,
                         r.Zlist, r.occ_d, r.planB, s);
            if (j == 0 && s < maxDump) { snprintf(name, sizeof name, "V_s%04d.c64", s); dumpDev(name, r.V_d, cb); }
            forwardPropagation(r.psi, r.V_d, r.frProp, r.t, p, r.params_d);
            if (j == 0 && s < maxDump) { snprintf(name, sizeof name, "psi_s%04d.c64", s); dumpDev(name, r.psi, cb); }
            if (j == 0 && s == 0) dumpDev("frProp.c64", r.frProp, cb);
        }
        snprintf(name, sizeof name, "psi_exit_cfg%03d.c64", j);
        dumpDev(name, r.psi, cb);
        cublasCaxpy(p->CU.cublasHandle, r.m12, &alpha, r.psi, 1, ew_d, 1);
        if (p->IM.mode == 0) {
            applyLensFunction(r.psi, k, p, r.params_d);
            intensityValues<<<p->CU.gS, p->CU.bS>>>(r.psi, r.m12);
        } else {
            diffractionPattern(r.psi, k, p, r.params_d);
        }
        cublasCaxpy(p->CU.cublasHandle, r.m12, &alpha, r.psi, 1, r.I_d, 1);
    }
    dumpDev("exitwave_avg.c64", ew_d, cb);
    dumpDev("I_d.c64", r.I_d, cb);
    addNoiseAndMtf(r.J_d, r.I_d, p->IM.pD, k, r.poissonState_d, p, r.params_d);
    dumpDev("J.f32", r.J_d, (size_t)p->IM.n1 * p->IM.n2 * sizeof(float));
    FILE* f = fopen((g_outdir + "/meta.txt").c_str(), "w");
    fprintf(f, "n1 %d\nn2 %d\nn3 %d\nm1 %d\nm2 %d\nm3 %d\nd1 %.9g\nd2 %.9g\nd3 %.9g\n"
               "lambda %.9g\nsigma %.9g\ngamma %.9g\nmode %d\nfrPh %d\nnAt %d\nnZ %d\n",
            p->IM.n1, p->IM.n2, p->IM.n3, p->IM.m1, p->IM.m2, p->IM.m3, p->IM.d1, p->IM.d2,
            p->IM.d3, p->EM.lambda, p->EM.sigma, p->EM.gamma, p->IM.mode, p->IM.frPh, r.nAt, r.nZ);
    fclose(f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "ref_harness: CUDA error %s\n", cudaGetErrorString(e)); return 4; }
    return 0;
}

static int modeTime(const std::string& in, int reps, int configs)
{
    Rig r; memset(&r, 0, sizeof r);
    if (!loadInput(in, &r.params, &r.Z_d, &r.xyz_d, &r.DWF_d, &r.occ_d)) return 2;
    rigSetup(r);
    params_t* p = r.params;
    const int k = 0;
    cublasScopy(p->CU.cublasHandle, r.nAt * 3, r.xyzTO_d, 1, r.xyz_d, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ms;
    for (int rep = 0; rep < reps + 1; rep++) {          /* rep 0 = warm-up */
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int j = 0; j < configs; j++) {
            incomingWave(r.psi, k, p, r.params_d);
            cublasScopy(p->CU.cublasHandle, r.nAt * 3, r.xyz_d, 1, r.xyzFP_d, 1);
            if (p->IM.frPh > 0) atomJitter(r.xyzFP_d, r.dwfState_d, r.nAt, r.DWF_d);
            for (int s = 0; s < p->IM.m3; s++) {
                phaseGrating(r.V_d, r.nAt, r.nZ, p, r.params_d, r.xyzFP_d, p->SAMPLE.imPot, r.Z_d,
                             r.Zlist, r.occ_d, r.planB, s);
                forwardPropagation(r.psi, r.V_d, r.frProp, r.t, p, r.params_d);
            }
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1);
        if (rep > 0) ms.push_back(t);
    }
    double sum = 0; for (float t : ms) sum += t;
    const double mean_ms = sum / ms.size();
    const double pxsl = (double)r.m12 * p->IM.m3 * configs;
    printf("{\"ref_time\": true, \"m1\": %d, \"m2\": %d, \"slices\": %d, \"configs\": %d, \"nAt\": %d, "
           "\"nZ\": %d, \"reps\": %d, \"ms_per_rep\": %.6f, \"mpx_slices_per_s\": %.6f}\n",
           p->IM.m1, p->IM.m2, p->IM.m3, configs, r.nAt, r.nZ, reps, mean_ms,
           pxsl / (mean_ms * 1e-3) / 1e6);
    return 0;
}

static int modeE2E(const std::string& in, int reps, int warm, double budget_s)
{
    /* atoms as a host [nAt][6] array, taken once from the file */
    params_t* p0 = NULL; int* Z_d = NULL; float *xyz_d = NULL, *DWF_d = NULL, *occ_d = NULL;
    confOption = 1;
    if (!getParams(in.c_str(), &p0, &Z_d, &xyz_d, &DWF_d, &occ_d)) return 2;
    const int nAt = p0->SAMPLE.nAt;
    std::vector<int> Zh(nAt); std::vector<float> xyz(3 * (size_t)nAt), dwf(nAt), occ(nAt);
    cudaMemcpy(Zh.data(), Z_d, nAt * sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(xyz.data(), xyz_d, 3 * (size_t)nAt * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(dwf.data(), DWF_d, nAt * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(occ.data(), occ_d, nAt * sizeof(float), cudaMemcpyDeviceToHost);
    std::vector<float> a6(6 * (size_t)nAt);
    for (int i = 0; i < nAt; i++) {
        a6[6 * i] = (float)Zh[i]; a6[6 * i + 1] = xyz[3 * i]; a6[6 * i + 2] = xyz[3 * i + 1];
        a6[6 * i + 3] = xyz[3 * i + 2]; a6[6 * i + 4] = dwf[i]; a6[6 * i + 5] = occ[i];
    }
    const int n123 = p0->IM.n1 * p0->IM.n2 * p0->IM.n3;
    const int m1 = p0->IM.m1, m2 = p0->IM.m2, frPh = p0->IM.frPh;
    std::vector<float> dst(n123);
    cudaFree(Z_d); cudaFree(xyz_d); cudaFree(DWF_d); cudaFree(occ_d);
    freeParams(&p0);
    std::vector<double> ms;
    int slices = 0;
    double spent_s = 0;
    for (int rep = 0; rep < reps + warm; rep++) {
        cudaDeviceSynchronize();
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        params_t* params = NULL; Z_d = NULL; xyz_d = DWF_d = occ_d = NULL;
        atomsFromExternal = true;
        printLevel = 0;
        if (!getParams(in.c_str(), &params, &Z_d, &xyz_d, &DWF_d, &occ_d)) return 2;
        readAtomsFromArray(params, &Z_d, &xyz_d, &DWF_d, &occ_d, a6.data(), nAt);
        g_e2e_dst = dst.data();
        buildMeasurements(params, Z_d, xyz_d, DWF_d, occ_d, (char*)"Measurements.bin", (char*)"results.emd");
        slices = params->IM.m3;
        freeParams(&params);
        cudaFree(xyz_d); cudaFree(DWF_d); cudaFree(occ_d); cudaFree(Z_d);
        cudaDeviceSynchronize();
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double call_ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
        if (rep >= warm) ms.push_back(call_ms);
        spent_s += call_ms * 1e-3;
        /* bounded run: stop once the time budget is used up (at least one timed call) */
        if (budget_s > 0 && spent_s > budget_s && !ms.empty()) break;
    }
    double sum = 0; for (double t : ms) sum += t;
    const double mean_ms = sum / ms.size();
    const int count = frPh > 0 ? frPh : 1;
    const double pxsl = (double)m1 * m2 * slices * count;
    double chk = 0; for (float v : dst) chk += v;
    printf("{\"ref_e2e\": true, \"m1\": %d, \"m2\": %d, \"slices\": %d, \"configs\": %d, \"nAt\": %d, "
           "\"reps\": %d, \"ms_per_call\": %.6f, \"mpx_slices_per_s\": %.6f, \"image_mean\": %.6f}\n",
           m1, m2, slices, count, nAt, (int)ms.size(), mean_ms, pxsl / (mean_ms * 1e-3) / 1e6, chk / n123);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: ref_harness run|trace|time <input> <outdir|reps> [arg]\n");
        return 1;
    }
    const std::string mode = argv[1];
    const std::string in = absPath(argv[2]);
    cudaSetDevice(gpu_index);
    if (mode == "time" || mode == "e2e") {
        /* side-effect files of the readers (dataFDES_used.cnf, ...) go to a scratch dir */
        const char* tmp = getenv("TMPDIR") ? getenv("TMPDIR") : "/tmp";
        if (chdir(tmp) != 0) return 3;
        if (mode == "e2e") return modeE2E(in, atoi(argv[3]), argc > 4 ? atoi(argv[4]) : 1, argc > 5 ? atof(argv[5]) : 0.0);
        return modeTime(in, atoi(argv[3]), argc > 4 ? atoi(argv[4]) : 1);
    }
    g_outdir = absPath(argv[3]);
    mkdir(g_outdir.c_str(), 0777);
    /* .qsc inputs reference their .cfg relative to the cwd: run from the input's directory,
     * unless it is read-only -- then the caller must have copied the inputs. */
    std::string dir = in.substr(0, in.find_last_of('/'));
    if (chdir(g_outdir.c_str()) != 0) return 3;
    if (strstr(in.c_str(), ".qsc") && chdir(dir.c_str()) != 0) return 3;
    if (mode == "run") {
        printLevel = argc > 4 ? atoi(argv[4]) : 2;
        params_t* params = NULL; int* Z_d = NULL; float *xyz_d = NULL, *DWF_d = NULL, *occ_d = NULL;
        if (!loadInput(in, &params, &Z_d, &xyz_d, &DWF_d, &occ_d)) return 2;
        std::string img = g_outdir + "/Measurements.bin", emd = g_outdir + "/results.emd";
        buildMeasurements(params, Z_d, xyz_d, DWF_d, occ_d, (char*)img.c_str(), (char*)emd.c_str());
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { fprintf(stderr, "ref_harness: CUDA error %s\n", cudaGetErrorString(e)); return 4; }
        return 0;
    }
    if (mode == "trace") return modeTrace(in, argc > 4 ? atoi(argv[4]) : 1 << 30);
    fprintf(stderr, "ref_harness: unknown mode %s\n", mode.c_str());
    return 1;
}
