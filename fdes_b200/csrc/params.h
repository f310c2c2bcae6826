// fdes_b200 -- host-side simulation parameters: the C++ mirror of the reference's params_t
// (include/paramStructure.h:48-162) and its .cnf reader / writer
// (src/paramStructure.cu:42-302, 360-487, 588-673, 1019-1077).
#pragma once
#include <string>
#include <vector>

namespace fdes {

// order of the aberration arrays: C1 A1 A2 B2 C3 A3 S3 A4 B4 D4 C5 A5 R5 S5
enum { AB_C1 = 0, AB_A1, AB_A2, AB_B2, AB_C3, AB_A3, AB_S3, AB_A4, AB_B4, AB_D4, AB_C5, AB_A5, AB_R5, AB_S5, AB_COUNT };
extern const char* const kAberrationNames[AB_COUNT];

struct Params {
    // constants (allocParams / defaultParams, src/paramStructure.cu:490-497, 696-700)
    float cst_m0 = 9.1093822e-31f, cst_c = 2.9979246e8f, cst_e = 1.6021766e-19f,
          cst_h = 6.6260696e-34f, cst_pi = 3.1415927f;
    // EM
    float E0 = 200e3f, gamma = 1.3913902f, lambda = 2.507934e-012f, sigma = 7288400.5f;
    float ab0[AB_COUNT] = {-6.1334e-008f, 0, 0, 0, 1e-3f, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    float ab1[AB_COUNT] = {0};
    float defocspread = 0.f, illangle = 0.f, mtfa = 1.f, mtfb = 0.f, mtfc = 0.f, mtfd = 0.f;
    float ObjAp = 11.1e-3f;
    // IM
    int mode = 0;
    int m1 = 4, m2 = 4, m3 = 1;
    float d1 = 0.25e-10f, d2 = 0.25e-10f, d3 = 2e-10f;
    int dn1 = 1, dn2 = 1, n1 = 2, n2 = 2, n3 = 1;
    int frPh = 0;
    float pD = 0.f, subSlTh = 2e-10f;
    std::vector<float> tiltspec, tiltbeam, defoci;   // 2*n3, 2*n3, n3
    float tilt_off[3] = {0.f, 0.f, 0.f};
    bool doBeamTilt = false;
    // SAMPLE / USER / COMMENT
    float imPot = 0.f;
    int nAt = 0;
    std::string sample_name = "Empty sample", material = "Nothing";
    std::string user_name = "John Smith", institution = "Europe University",
                department = "Electron Microscopy Facility", email = "john.smith@uni.eu";
    std::string comments = "This is FDES's default comment";
};

struct Atoms {
    std::vector<int> Z;
    std::vector<float> xyz;   // [nAt][3], metres
    std::vector<float> dwf;   // m^2
    std::vector<float> occ;
    int size() const { return (int)Z.size(); }
};

// getParams (src/paramStructure.cu:588-635): defaults, .cnf keys, consitentParams.  When
// atoms_from_external is set the `atom:` lines are ignored (src/paramStructure.cu:268).
// Returns false if the file cannot be opened.
bool read_cnf(const char* file, Params& p, Atoms* atoms, bool atoms_from_external);
// readQsc (src/rwQsc.cu:8-1088) with the QSTEM .cfg unit-cell reader behind it
// (qstem-libs/fileio_fftw3.cpp:721-776, 908-975, 1188-1657); see qsc.cpp for what is kept and what
// is refused.  Throws std::runtime_error on unusable input, returns false if the file cannot be opened.
bool read_qsc(const char* file, Params& p, Atoms* atoms, bool atoms_from_external);
// Reader selected by the file name like src/FDESExport.cu:85-102 (.cnf, else .qsc)
bool read_input(const char* file, Params& p, Atoms* atoms, bool atoms_from_external);
bool is_qsc_name(const char* file);
// consitentParams (src/paramStructure.cu:637-673)
void consistent_params(Params& p);
// readAtomsFromArray (src/paramStructure.cu:304-345): [numAtoms][6] = Z x y z DWF occ; occupancy is
// truncated to an integer exactly like the reference (:323).
void atoms_from_array(const float* atomsArray, int numAtoms, Atoms& atoms);
// writeConfig (src/paramStructure.cu:360-487)
bool write_cnf(const char* file, const Params& p, const Atoms& atoms, int gpu_index);
// subSliceRatio / setSubSlices (src/crystalMaker.cu:720-743)
float sub_slice_ratio(float slice, float subSlice);
void set_sub_slices(Params& p, float ratio);
// listOfElements (src/crystalMaker.cu:539-570): first-appearance order
std::vector<int> list_of_elements(const std::vector<int>& Z);
// writeBinary (src/rwBinary.cpp): raw float32
bool write_binary(const char* file, const float* data, size_t n);

}  // namespace fdes
