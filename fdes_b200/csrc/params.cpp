// fdes_b200 -- .cnf reader/writer and derived parameters (host, C++).
// Behavioural mirror of the reference reader so the same input file yields the same params_t
// and the same atom list, including its observable quirks:
//   * lines are consumed with a 100-byte fgets for parameters (src/paramStructure.cu:46,66) and a
//     200-byte fgets for atoms (:1023,1051);
//   * the field name is only refreshed when sscanf("%s") converts, so a blank line (or the
//     failed read at end-of-file) re-dispatches the previous field name: the specimen_tilt /
//     beam_tilt / defoci counters advance on blank lines, and a file that ends with
//     "atom: ...\n" yields its last atom twice (numberOfAtoms / readCoordinates, :1019-1077).
#include "params.h"
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <charconv>
#include <string>
#include <thread>
#include <vector>

namespace fdes {

const char* const kAberrationNames[AB_COUNT] = {"C1", "A1", "A2", "B2", "C3", "A3", "S3",
                                                "A4", "B4", "D4", "C5", "A5", "R5", "S5"};

namespace {

struct FloatKey { const char* key; int cmp; float Params::*dst; };
struct IntKey { const char* key; int cmp; int Params::*dst; };
struct StrKey { const char* key; int cmp; const char* fmt; std::string Params::*dst; };

const FloatKey kFloatKeys[] = {
    {"voltage:", 8, &Params::E0},
    {"focus_spread:", 12, &Params::defocspread},
    {"illumination_angle:", 19, &Params::illangle},
    {"mtf_a:", 6, &Params::mtfa},
    {"mtf_b:", 6, &Params::mtfb},
    {"mtf_c:", 6, &Params::mtfc},
    {"mtf_d:", 6, &Params::mtfd},
    {"objective_aperture:", 19, &Params::ObjAp},
    {"pixel_size_x:", 13, &Params::d1},
    {"pixel_size_y:", 13, &Params::d2},
    {"pixel_size_z:", 13, &Params::d3},
    {"absorptive_potential_factor:", 28, &Params::imPot},
    {"pixel_dose:", 11, &Params::pD},
    {"subpixel_size_z:", 16, &Params::subSlTh},
};
const IntKey kIntKeys[] = {
    {"sample_size_x:", 14, &Params::m1}, {"sample_size_y:", 14, &Params::m2},
    {"sample_size_z:", 14, &Params::m3}, {"border_size_x:", 14, &Params::dn1},
    {"border_size_y:", 14, &Params::dn2}, {"image_size_x:", 13, &Params::n1},
    {"image_size_y:", 13, &Params::n2},  {"image_size_z:", 13, &Params::n3},
    {"frozen_phonons:", 15, &Params::frPh}, {"mode:", 5, &Params::mode},
};
const StrKey kStrKeys[] = {
    {"user_name:", 11, "user_name: %1023[^\n]", &Params::user_name},
    {"institution:", 12, "institution: %1023[^\n]", &Params::institution},
    {"department:", 11, "department: %1023[^\n]", &Params::department},
    {"email:", 6, "email: %1023[^\n]", &Params::email},
    {"comment:", 8, "comment: %1023[^\n]", &Params::comments},
    {"sample_name:", 12, "sample_name: %1023[^\n]", &Params::sample_name},
    {"material:", 9, "material: %1023[^\n]", &Params::material},
};

void dispatch_line(const char* field, const char* line, Params& p, int& ts_i, int& tb_i, int& df_i)
{
    // atom lines carry no parameter key (none of the keys below starts with "atom:"); skipping
    // them up front keeps large specimen files from costing ~60 string compares per atom
    if (!strncmp(field, "atom:", 5)) return;
    for (const FloatKey& k : kFloatKeys)
        if (!strncmp(field, k.key, k.cmp)) sscanf(line, "%*s %g", &(p.*(k.dst)));
    for (int a = 0; a < AB_COUNT; a++) {
        char key[4] = {kAberrationNames[a][0], kAberrationNames[a][1], ':', 0};
        if (!strncmp(field, key, 3)) {
            if (a == AB_C1 || a == AB_C3 || a == AB_C5) sscanf(line, "%*s %g", &p.ab0[a]);
            else sscanf(line, "%*s %g %g", &p.ab0[a], &p.ab1[a]);
        }
    }
    for (const IntKey& k : kIntKeys)
        if (!strncmp(field, k.key, k.cmp)) sscanf(line, "%*s %i", &(p.*(k.dst)));
    for (const StrKey& k : kStrKeys)
        if (!strncmp(field, k.key, k.cmp)) {
            char buf[1024];
            if (sscanf(line, k.fmt, buf) == 1) p.*(k.dst) = buf;
        }
    if (!strncmp(field, "specimen_tilt_offset_x:", 23)) sscanf(line, "%*s %g", &p.tilt_off[0]);
    if (!strncmp(field, "specimen_tilt_offset_y:", 23)) sscanf(line, "%*s %g", &p.tilt_off[1]);
    if (!strncmp(field, "specimen_tilt_offset_z:", 23)) sscanf(line, "%*s %g", &p.tilt_off[2]);
    auto grow = [](std::vector<float>& v, size_t n) { if (v.size() < n) v.resize(n, 0.f); };
    if (!strncmp(field, "specimen_tilt:", 14)) {
        grow(p.tiltspec, 2 * (size_t)ts_i + 2);
        sscanf(line, "%*s %g %g", &p.tiltspec[2 * ts_i], &p.tiltspec[2 * ts_i + 1]);
        ts_i += 1;
    }
    if (!strncmp(field, "beam_tilt:", 10)) {
        grow(p.tiltbeam, 2 * (size_t)tb_i + 2);
        sscanf(line, "%*s %g %g", &p.tiltbeam[2 * tb_i], &p.tiltbeam[2 * tb_i + 1]);
        tb_i += 1;
    }
    if (!strncmp(field, "defoci:", 7)) {
        grow(p.defoci, (size_t)df_i + 1);
        sscanf(line, "%*s %g", &p.defoci[df_i]);
        df_i += 1;
    }
}

void read_atoms(FILE* fr, Atoms& atoms)
{
    fseek(fr, 0, SEEK_SET);
    char line[200], field[200] = "";
    atoms = Atoms();
    while (!feof(fr)) {
        if (fgets(line, sizeof line, fr) != NULL) sscanf(line, "%199s", field);
        if (!strncmp(field, "atom:", 5)) {
            int Z = 0;
            float v[5] = {0, 0, 0, 0, 0};
            sscanf(line, "%*s %i %g %g %g %g %g", &Z, &v[0], &v[1], &v[2], &v[3], &v[4]);
            atoms.Z.push_back(Z);
            atoms.xyz.push_back(v[0]); atoms.xyz.push_back(v[1]); atoms.xyz.push_back(v[2]);
            atoms.dwf.push_back(v[3]);
            atoms.occ.push_back(v[4]);
        }
        line[0] = '#';   // resetLine: a stale buffer re-parses as "#tom: ..." (same numbers)
    }
}

}  // namespace

void consistent_params(Params& p)
{
    const float E0 = p.E0;
    const float m0 = 9.1093822f, c = 2.9979246f, e = 1.6021766f, h = 6.6260696f;
    const float pi = p.cst_pi;
    p.gamma = 1.f + E0 * e / m0 / c / c * 1e-4f;
    p.lambda = h / sqrtf(2.f * m0 * e) * 1e-9f / sqrtf(E0 * (1.f + E0 * e / 2.f / m0 / c / c * 1e-4f));
    p.sigma = 2.f * pi * p.gamma * p.lambda * m0 * e / h / h * 1e18f;
    p.m1 = p.n1 + 2 * p.dn1;
    p.m2 = p.n2 + 2 * p.dn2;
    p.tiltspec.resize(2 * (size_t)p.n3, 0.f);
    p.tiltbeam.resize(2 * (size_t)p.n3, 0.f);
    p.defoci.resize((size_t)p.n3, 0.f);
    float flag = 0.f;
    for (int j = 0; j < p.n3 * 2; j++) flag += fabsf(p.tiltbeam[j]);
    p.doBeamTilt = !(flag < (FLT_MIN * ((float)p.n3 * 2)));
}

bool read_cnf(const char* file, Params& p, Atoms* atoms, bool atoms_from_external)
{
    FILE* fr = fopen(file, "rt");
    if (fr == NULL) {
        fprintf(stderr, "\n  Not able to read simulation configuration from %s\n", file);
        return false;
    }
    p = Params();
    p.cst_pi = 3.141592654f;   // allocParams value (same float32)
    p.n3 = 1000;               // getParams reads into a 1000-entry scratch params first
    int ts_i = 0, tb_i = 0, df_i = 0;
    char line[100], field[100] = "";
    do {
        if (fgets(line, sizeof line, fr) != NULL) {
            // sscanf(line, "%99s", field): first whitespace-delimited token, field untouched when
            // the line has none (hand-rolled: a specimen file is mostly atom lines)
            const char* s = line;
            while (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\v' || *s == '\f' || *s == '\r') s++;
            if (*s) {
                int n = 0;
                while (*s && !(*s == ' ' || *s == '\t' || *s == '\n' || *s == '\v' || *s == '\f' || *s == '\r') && n < 99) field[n++] = *s++;
                field[n] = 0;
            }
        }
        dispatch_line(field, line, p, ts_i, tb_i, df_i);
    } while (!feof(fr));
    if (p.n3 < 1 || p.n3 > 1000) {
        fprintf(stderr, "  image_size_z must be in [1, 1000] (got %d)\n", p.n3);
        fclose(fr);
        return false;
    }
    if (!atoms_from_external && atoms) {
        read_atoms(fr, *atoms);
        p.nAt = atoms->size();
        fprintf(stderr, "  Number of atoms in the specimen: %i \n ", p.nAt);
    }
    fclose(fr);
    consistent_params(p);
    return true;
}

void atoms_from_array(const float* a, int numAtoms, Atoms& atoms)
{
    atoms = Atoms();
    atoms.Z.resize(numAtoms); atoms.xyz.resize(3 * (size_t)numAtoms);
    atoms.dwf.resize(numAtoms); atoms.occ.resize(numAtoms);
    for (int i = 0; i < numAtoms; i++) {
        atoms.Z[i] = (int)a[6 * i + 0];
        atoms.xyz[3 * i + 0] = a[6 * i + 1];
        atoms.xyz[3 * i + 1] = a[6 * i + 2];
        atoms.xyz[3 * i + 2] = a[6 * i + 3];
        atoms.dwf[i] = a[6 * i + 4];
        atoms.occ[i] = (float)(int)a[6 * i + 5];
    }
}

float sub_slice_ratio(float slice, float subSlice)
{
    float ratio = 1.f;
    if ((subSlice > 1e-12f) && (subSlice < slice)) ratio = ceilf(slice / subSlice);
    return ratio;
}

void set_sub_slices(Params& p, float ratio)
{
    p.m3 = (int)(((float)p.m3) * ratio);
    p.d3 /= ratio;
}

std::vector<int> list_of_elements(const std::vector<int>& Z)
{
    std::vector<int> out;
    for (int z : Z) {
        bool seen = false;
        for (int q : out) if (q == z) seen = true;
        if (!seen) out.push_back(z);
    }
    return out;
}

bool write_binary(const char* file, const float* data, size_t n)
{
    FILE* f = fopen(file, "wb");
    if (!f) { fprintf(stderr, "  Cannot open %s for writing\n", file); return false; }
    const size_t w = fwrite(data, sizeof(float), n, f);
    fclose(f);
    return w == n;
}

bool write_cnf(const char* file, const Params& p, const Atoms& atoms, int gpu_index)
{
    FILE* fw = fopen(file, "wt");
    if (!fw) return false;
    auto F = [&](const char* k, float v, const char* c) { fprintf(fw, "%s %14.8g %s\n", k, v, c); };
    auto I = [&](const char* k, int v, const char* c) { fprintf(fw, "%s %i %s\n", k, v, c); };
    fprintf(fw, "# Parameter file. Let Comments be preceded by '#'\n");
    fprintf(fw, "\n# Nature's constants\n#-------------------\n\n");
    F("m0: ", p.cst_m0, " # Electron rest mass [kg]");
    F("c: ", p.cst_c, " # Speed of light [m/s]");
    F("e: ", p.cst_e, " # Elementary charge [C]");
    F("h: ", p.cst_h, " # Planck's constant [Js]");
    F("pi: ", p.cst_pi, " # Pi [dimensionless]");
    fprintf(fw, "\n# Microscope parameters\n#----------------------\n\n");
    F("voltage: ", p.E0, " # Acceleration voltage [V]");
    F("gamma: ", p.gamma, " # From relativity: 1+e*E0/m0/c^2");
    F("lambda: ", p.lambda, " # Electron wavelength [m]");
    F("sigma: ", p.sigma, " # Interaction constant [1/(Vm)]");
    fprintf(fw, "\n");
    F("focus_spread: ", p.defocspread, " # Defocus spread for the temporal partial coherence [m]");
    F("illumination_angle: ", p.illangle, " # illumination half angle characterizing the spatial partial coherence [rad]");
    F("mtf_a: ", p.mtfa, " # MTF parameters, see Microsc. Microanal. 18 (2012) 336-342.");
    F("mtf_b: ", p.mtfb, "");
    F("mtf_c: ", p.mtfc, "");
    F("mtf_d: ", p.mtfd, "");
    F("objective_aperture: ", p.ObjAp, " # Radius of the objective aperture [rad]");
    fprintf(fw, "\n# aberration coefficients\n");
    for (int a = 0; a < AB_COUNT; a++)
        fprintf(fw, "%s:  %14.8g %14.8g\n", kAberrationNames[a], p.ab0[a], p.ab1[a]);
    fprintf(fw, "\n# Imaging parameters\n#-------------------\n\n");
    I("mode: ", p.mode, " # 0, 1 or 2 for imaging, diffraction or CBED, resp.");
    I("gpu_index: ", gpu_index, " # Number of the device the code runs on, it's often 0 or 1");
    I("sample_size_x: ", p.m1, " # Width of the object: length of 2nd dimension or number of columns");
    I("sample_size_y: ", p.m2, " # Height of the object: length of 1st dimension or number rows");
    I("sample_size_z: ", p.m3, " # Depth of the object: number of sample_size_y by sample_size_x slices");
    F("pixel_size_x: ", p.d1, " # Length of 2nd dimension of the voxels [m]");
    F("pixel_size_y: ", p.d2, " # Length of 1st dimension of the voxels [m]");
    F("pixel_size_z: ", p.d3, " # Length of 3rd dimension of the voxels [m] of the saved potential_slices");
    I("border_size_x: ", p.dn1, " # Padding, sample_size_x = image_size_x + 2*border_size_x");
    I("border_size_y: ", p.dn2, " # Padding, sample_size_y = image_size_y + 2*border_size_y");
    I("image_size_x: ", p.n1, " # Width of the measurements: length of 2nd dimension or number of columns");
    I("image_size_y: ", p.n2, " # Height of the measurements: length of 1st dimension or number rows");
    I("image_size_z: ", p.n3, " # Number of image_size_y by image_size_x measurements");
    F("specimen_tilt_offset_x: ", p.tilt_off[0], " # Initial tilt of the object around the second axis [rad]");
    F("specimen_tilt_offset_y: ", p.tilt_off[1], " # Initial tilt of the object around the first axis [rad]");
    F("specimen_tilt_offset_z: ", p.tilt_off[2], " # Initial tilt of the object around the third axis [rad]");
    fprintf(fw, "%s %d %s\n", "frozen_phonons: ", p.frPh, " # Number of frozen phonon iterations, set to 0 if none are needed");
    F("pixel_dose: ", p.pD, " # mean number of electrons per pixel, set to zero for noise-free imaging");
    F("subpixel_size_z: ", p.subSlTh, " # The images are calculated with (approx.!) this slice thickness [m]");
    fprintf(fw, "\n# Sample properties\n#------------------\n\n");
    F("absorptive_potential_factor: ", p.imPot, " # Imaginary potential factor to approximate absorption: V <- V + iV * absorptive_potential_factor");
    fprintf(fw, "\n# Specimen tilts, beam tilts and defoci\n#--------------------------------------\n\n");
    fprintf(fw, "# Tilts of the specimen [rad]. Quantities in each row:\n");
    fprintf(fw, "# specimen_tilt_x,  specimen_tilt_y. Tilts around the second and the first specimen axis resp.\n");
    for (int i = 0; i < p.n3; i++) fprintf(fw, "%s %14.8g %14.8g\n", "specimen_tilt: ", p.tiltspec[2 * i], p.tiltspec[2 * i + 1]);
    fprintf(fw, "\n# Tilts of the beam [rad].  Quantities in each row:\n");
    fprintf(fw, "# beam_tilt_x,  beam_tilt_y. Tilts around the second and the first specimen axis resp.\n");
    for (int i = 0; i < p.n3; i++) fprintf(fw, "%s %14.8g %14.8g\n", "beam_tilt: ", p.tiltbeam[2 * i], p.tiltbeam[2 * i + 1]);
    fprintf(fw, "\n# Defoci values [m]\n");
    for (int i = 0; i < p.n3; i++) fprintf(fw, "%s %14.8g\n", "defoci: ", p.defoci[i]);
    fprintf(fw, "\n# The atoms in the sample\n#------------------------\n\n");
    fprintf(fw, "Number of atoms: %d\n", atoms.size());
    fprintf(fw, "\n# List of atoms. Quantities in each row:\n");
    fprintf(fw, "# Atomic no.; x-, y- and z-coordinate [m]; Debeye-Waller factor [m^2]; occupancy.\n");
    // The atom table ("%i %14.8g %14.8g %14.8g %14.8g %14.8g \n" in the reference, src/paramStructure.cu:480-484) is
    // formatted with std::to_chars -- specified to give the digits of printf's %.8g in the C locale, at a fraction
    // of snprintf's cost -- by at most four threads into memory and written in one go: a specimen of 10^4 atoms is
    // 6*10^4 conversions per call, which with snprintf took longer than the simulation itself and, with one
    // process per GPU, saturated the host cores of an 8-GPU box.
    const int nAt = atoms.size();
    const int nthreads = std::max(1, std::min(4, nAt / 2048));
    std::vector<std::string> chunks(nthreads);
    auto put_g = [](char* out, float v) -> char* {     // "%14.8g": right-aligned in 14 columns (or wider)
        char tmp[32];
        const auto r = std::to_chars(tmp, tmp + sizeof tmp, v, std::chars_format::general, 8);
        const int n = (int)(r.ptr - tmp);
        for (int i = n; i < 14; i++) *out++ = ' ';
        memcpy(out, tmp, (size_t)n);
        return out + n;
    };
    auto format_range = [&](int t) {
        const int lo = (int)((long long)nAt * t / nthreads), hi = (int)((long long)nAt * (t + 1) / nthreads);
        std::string& s = chunks[t];
        s.reserve((size_t)(hi - lo) * 90);
        char buf[200];
        for (int j = lo; j < hi; j++) {
            char* q = std::to_chars(buf, buf + 16, atoms.Z[j]).ptr;
            const float v[5] = {atoms.xyz[3 * j + 0], atoms.xyz[3 * j + 1], atoms.xyz[3 * j + 2], atoms.dwf[j], atoms.occ[j]};
            for (int k = 0; k < 5; k++) { *q++ = ' '; q = put_g(q, v[k]); }
            *q++ = ' '; *q++ = '\n';
            s.append(buf, (size_t)(q - buf));
        }
    };
    std::vector<std::thread> workers;
    for (int t = 1; t < nthreads; t++) workers.emplace_back(format_range, t);
    format_range(0);
    for (auto& w : workers) w.join();
    for (const std::string& s : chunks) fwrite(s.data(), 1, s.size(), fw);
    fclose(fw);
    return true;
}

}  // namespace fdes
