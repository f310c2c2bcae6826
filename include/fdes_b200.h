/* fdes_b200 -- C ABI of the B200-native FDES forward-multislice library (libfdes_b200.so).
 *
 * Plain C: pointers and sizes only, no C++ or torch types.  Two layers:
 *
 *  (1) The drop-in symbol.  `FDES` has the exact signature the reference exports from
 *      libFDES_SHARED_LIB.so (reference src/FDESExport.cu:59-60) and that
 *      Python/pyFDES.py:33-41 binds through ctypes, so `ctypes.CDLL("libfdes_b200.so").FDES`
 *      (or a symlink named libFDES_SHARED_LIB.so) replaces it without touching the caller.
 *
 *  (2) A session API over the same engine.  It exposes the seam the reference keeps internal
 *      (buildMeasurements, include/crystalMaker.h:77) so a host framework can shard frozen-phonon
 *      configurations over several GPUs (one process per GPU) and reduce the partial intensity /
 *      exit-wave sums itself (e.g. torch.distributed all_reduce over NCCL), plus the building
 *      blocks the parity tests compare against the oracle.
 *
 * All functions returning int return 0 on success and -1 on failure; the message is available
 * from fdes_b200_last_error().  There is no CPU fallback: without a CUDA device every entry point
 * that computes fails.
 */
#ifndef FDES_B200_H
#define FDES_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- (1) drop-in for the reference export ------------------------------------------------
 * Replaces: extern "C" void FDES(...)                     reference src/FDESExport.cu:59-178
 *   gpu_Index      CUDA device ordinal                    (:68, :106)
 *   print_Level    0 images | 1 + potential | 2 + exit waves (:69)
 *   input_name     parameter file: ".emd", ".cnf" or ".qsc", tested in that order     (:73-102)
 *   image_name     raw float32 output file [n3][n2][n1]    (src/crystalMaker.cu:399)
 *   emd_save_name  results file, EMD/HDF5 with the reference's groups (src/rwHdf5.cu:27-1084)
 *   atomsArray     host float32 [numAtoms][6] = Z, x, y, z [m], DWF [m^2], occupancy
 *                                                          (src/paramStructure.cu:316-324)
 *   dstImage       host float32 [n3][n2][n1], caller-allocated (:162, src/paramStructure.cu:347-359)
 * Like the reference it returns void and terminates the process (exit) on unusable input. */
void FDES(int gpu_Index, int print_Level, char* input_name, char* image_name, char* emd_save_name,
          float* atomsArray, int numAtoms, float* dstImage);

/* ---- (2) session API ---------------------------------------------------------------------- */
typedef struct fdes_b200_sim fdes_b200_sim;

const char* fdes_b200_last_error(void);
/* Device memory blocks are cached per process between simulations (FDES() calls included) so that
 * repeated calls do not pay cudaMalloc/cudaFree; this returns the idle ones to the driver. */
void fdes_b200_release_cache(void);
int fdes_b200_version(void);

/* Host-only (no CUDA call): parse a .cnf (or .qsc / .emd) exactly as fdes_b200_open_cnf does and report what the
 * engine would run (readConfig + consitentParams, src/paramStructure.cu:42-302, 637-673; sub-slice
 * logic src/crystalMaker.cu:720-743).  dims[10] = n1 n2 n3 m1 m2 m3(after sub-slicing) nAt nZ
 * frozen_phonons mode; scalars[8] as fdes_b200_get_scalars; per_k [n3][5] = specimen_tilt x y,
 * beam_tilt x y, defocus (may be NULL); atoms6_out [max_atoms][6] (may be NULL).
 * Returns the number of atoms in the file, -1 on failure. */
int fdes_b200_parse_cnf(const char* cnf_path, int* dims, float* scalars, float* per_k,
                        float* atoms6_out, int max_atoms);

/* Host-only (no CUDA call): read a .cnf / .qsc / .emd parameter file exactly as fdes_b200_open_cnf does and
 * write parameters + atoms back in .cnf syntax -- the side-effect file the reference leaves behind
 * (writeConfig, src/paramStructure.cu:360-487: "dataFDES_used.cnf" from getParams :629-631,
 * "ParamsUsedQsc.txt" from readQsc, src/rwQsc.cu:1084). */
int fdes_b200_write_used_cnf(const char* input_path, const char* out_path);

/* Host-only (no CUDA call): write an EMD (HDF5) file with the reference's layout (writeHdf5,
 * src/rwHdf5.cu:27-1084; the bytes are produced by the library's own serialiser, there is no libhdf5
 * here) from the parameters + atoms of input_path (.cnf / .qsc / .emd) and the caller's arrays:
 * image_host [n3][n2][n1] -> /data/images/data [n1][n2][n3]; potential_host [pot_slices][m2][m1][2]
 * -> /data/potential_slices/data [m1][m2][pot_slices][2]; exitwave_host [n3][m2][m1][2] ->
 * /data/exit_wave/data [m1][m2][n3][2].  Any array may be NULL (all NULL = the "config.emd" of the
 * 6-argument writeHdf5, :1085-1945).  FDES() writes emd_save_name through the same code. */
int fdes_b200_write_emd(const char* input_path, const char* emd_path, const float* image_host,
                        const float* potential_host, int pot_slices, const float* exitwave_host);

/* Open a simulation from a parameter file, selected by name like src/FDESExport.cu:85-102:
 * .emd (readHdf5, src/rwHdf5.cu:1946-2571, parsed by the library's own HDF5 reader),
 * .cnf (reader: getParams, reference src/paramStructure.cu:588-635) or QSTEM .qsc + the .cfg unit
 * cell it names (readQsc, src/rwQsc.cu:8-1088; fdes_b200/csrc/qsc.cpp lists what is refused).
 * atoms6 == NULL: atoms come from the file's `atom:` lines; otherwise [numAtoms][6] like FDES().
 * batch: phonon configurations advanced together (0 = automatic).
 * rank/world: this process handles configurations [count*rank/world, count*(rank+1)/world).
 * want_exitwave: keep the coherent exit-wave average (reference print_level 2). */
fdes_b200_sim* fdes_b200_open_cnf(const char* cnf_path, const float* atoms6, int numAtoms,
                                  int gpu_index, int batch, int rank, int world, int want_exitwave);
/* The same session on `ngpus` devices of THIS process (one engine and, while they compute, one host
 * thread per device) -- the multi-GPU form of the seam below buildMeasurements
 * (include/crystalMaker.h:77; the reference itself is single-GPU, cudaSetDevice at src/FDES.cu:165).
 * What is sharded follows the reference's loop nest (src/crystalMaker.cu:324-372): with frozen phonons
 * the configurations j of every measurement k (contiguous blocks per device, identical RNG streams +
 * burn-in, partial intensities / exit waves summed onto gpu_indices[0] by a kernel that reads the
 * peers' buffers over NVLink, detector tail once per k); without them the measurements k of a tilt /
 * defocus series, or the probe positions of fdes_b200_stem_scan.  fdes_b200_simulate,
 * fdes_b200_stem_scan and fdes_b200_close work on such a session unchanged; the building blocks further down act on
 * the engine of gpu_indices[0].  FDES() opens its session this way when the environment variable
 * FDES_B200_GPUS is set ("4" = gpu_Index .. gpu_Index + 3, or a list "0,2,5"); the CLI has --gpus. */
fdes_b200_sim* fdes_b200_open_multi(const char* cnf_path, const float* atoms6, int numAtoms,
                                    const int* gpu_indices, int ngpus, int batch, int want_exitwave);
/* engines (GPUs) the session actually uses: min(ngpus, independent units) */
int fdes_b200_num_gpus(const fdes_b200_sim* sim);
/* the FDES_B200_GPUS / --gpus syntax -> device ordinals; returns how many (out may be NULL when max_out = 0) */
int fdes_b200_parse_gpu_list(const char* spec, int first, int* out, int max_out);
void fdes_b200_close(fdes_b200_sim* sim);

/* dims[10] = n1 n2 n3 m1 m2 m3(after sub-slicing) nAt nZ phonon_configs batch */
int fdes_b200_get_dims(const fdes_b200_sim* sim, int* dims);
/* scalars[8] = lambda sigma gamma d1 d2 d3(after sub-slicing) E0 imPot */
int fdes_b200_get_scalars(const fdes_b200_sim* sim, float* scalars);

/* Device accumulators for the partial sums of this rank (float32 on the sim's device):
 * intensity_dev [m2*m1], exitwave_dev [m2*m1*2] (may be NULL).  NULL restores the internal ones. */
int fdes_b200_set_accumulators(fdes_b200_sim* sim, float* intensity_dev, float* exitwave_dev);
/* All phonon configurations of this rank for measurement k (loop body of
 * src/crystalMaker.cu:324-367): accumulators <- sum_j (.)/count. */
int fdes_b200_run_k(fdes_b200_sim* sim, int k);
/* Detector tail on the (reduced) accumulators (addNoiseAndMtf, src/crystalMaker.cu:579-613):
 * image_host [n2*n1]; exitwave_host [m2*m1*2] or NULL. */
int fdes_b200_finish_k(fdes_b200_sim* sim, int k, float* image_host, float* exitwave_host);
/* Whole single-GPU run: image_host [n3][n2][n1], exitwave_host [n3][m2][m1][2] or NULL.
 * Host buffers in, host buffers out (this is what FDES() calls). */
int fdes_b200_simulate(fdes_b200_sim* sim, float* image_host, float* exitwave_host);
/* Untilted phonon-free potential slices, original slicing: out_host [m3_orig][m2][m1][2]
 * (src/crystalMaker.cu:381-397).  m3_orig is returned by fdes_b200_potential_slices_count. */
int fdes_b200_potential_slices_count(const fdes_b200_sim* sim);
int fdes_b200_potential(fdes_b200_sim* sim, float* out_host);

/* ---- STEM probe scan (extension; FDES has mode 2 = one CBED probe at the grid centre only,
 * src/multisliceSimulation.cu:572-581, and rejects mode: STEM inputs, src/rwQsc.cu:33-44) --------
 * The .cnf must select mode 2.  A scan position is the reference's probe shifted periodically to
 * xy_host[i] = (x, y) [m] relative to the grid centre; the specimen stays fixed.  Detector d
 * integrates the diffraction intensity |FFT2 psi|^2 / N^2 (diffractionPattern,
 * src/crystalMaker.cu:700-718) over k_in^2 <= |k|^2 < k_out^2 with k = sin(theta * 1e-3) / lambda,
 * theta = det_mrad_host[d] = (inner, outer) [mrad] (detector convention of src/rwQsc.cu:723-730).
 * out_host [nprobes][ndet], averaged over the frozen-phonon configurations; an empty specimen
 * gives n1*n2 on a detector that covers the whole pattern.  Returns the CUDA-event milliseconds
 * of the scan (negative on failure). */
double fdes_b200_stem_scan(fdes_b200_sim* sim, int k, int nprobes, const float* xy_host, int ndet,
                           const float* det_mrad_host, float* out_host);

/* Host-only: the scan raster and detectors a QSTEM `mode: STEM` .qsc describes -- the keys readQsc
 * parses (scan_x_start/stop/pixels, scan_y_*, src/rwQsc.cu:444-466; `detector: inner outer name ..`
 * [mrad], :698-735) but FDES never uses -- in the form fdes_b200_stem_scan takes: nxy[2] = pixels in
 * x, y; xy_host [nx][ny][2] probe positions [m] (QSTEM raster start + i (stop - start) / pixels, moved
 * into the frame of the atoms as readQsc centres them); det_mrad_host [ndet][2].  Returns the number
 * of detectors, -1 on failure.  Call once with NULL arrays to size them. */
int fdes_b200_qsc_scan(const char* qsc_path, int* nxy, float* xy_host, int max_probes, float* det_mrad_host,
                       int max_det);

/* ---- building blocks (parity tests, benchmarks) ------------------------------------------- */
/* next frozen-phonon coordinates for measurement k -> xyz_host [nAt][3]
 * (atomJitter_d, src/crystalMaker.cu:37-48; advances the XORWOW streams) */
int fdes_b200_jitter_next(fdes_b200_sim* sim, int k, float* xyz_host);
/* integer bin tuples (i1, i2, i3, species index) per atom, -1 when the atom is rejected
 * (squareAtoms_d index arithmetic, src/crystalMaker.cu:85-92): bins_host [nAt][4] */
int fdes_b200_bin_atoms(fdes_b200_sim* sim, const float* xyz_host, int* bins_host);
/* phase grating of one slice (phaseGrating, src/crystalMaker.cu:507-536): V_host [m2*m1*2] */
int fdes_b200_phase_grating(fdes_b200_sim* sim, const float* xyz_host, int slice, float* V_host);
/* exit wave of one configuration with given coordinates (incomingWave + m3 x
 * (phaseGrating + forwardPropagation)): psi_host [m2*m1*2] */
int fdes_b200_exit_wave(fdes_b200_sim* sim, const float* xyz_host, int k, float* psi_host);
/* Throughput loop: `configs` configurations with everything resident in HBM; returns the
 * CUDA-event milliseconds of the loop (negative on failure). */
double fdes_b200_bench_configs(fdes_b200_sim* sim, int k, int configs);
/* Average launch duration [ms] of each of the six per-slice sweeps (S1 density rows, S2 potential
 * columns, S3 transmission rows, S4 band-limit columns, S5 multiply rows, S6 propagate columns)
 * over `reps` back-to-back launches on a prepared batch -- the live kernel times behind the
 * roofline figures of bench.py. */
int fdes_b200_time_sweeps(fdes_b200_sim* sim, int k, int batch, int reps, float* ms6);
/* counters[4] = slices executed, kernels launched (since open or last reset), band columns
 * (columns kept by the 2/3 limit, rounded to tiles), 0 */
int fdes_b200_get_counters(fdes_b200_sim* sim, long long* counters, int reset);
/* 2-D complex64 FFT of a host array [N][N] in place with the library's own sweeps
 * (dir -1 forward, +1 unnormalised inverse) -- replaces cufftExecC2C for the tests. */
int fdes_b200_fft2d(float* data_host, int N, int dir, int gpu_index);
/* stable radix sort + row pointers on host arrays (test hook for the binning pipeline):
 * keys [n] (values < nkeys are valid), cols [n], w [n] sorted in place; rowptr [nkeys+1]. */
int fdes_b200_sort_records(unsigned int* keys, int* cols, float* w, int n, int nkeys, int* rowptr,
                           int gpu_index);

#ifdef __cplusplus
}
#endif
#endif /* FDES_B200_H */
