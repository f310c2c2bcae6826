// fdes_b200 -- reader for QSTEM .qsc parameter files and the .cfg unit-cell files they name.
//
// Mirrors what the reference does for `--input_name f.qsc`: readQsc (src/rwQsc.cu:8-1088) on top of
// the QSTEM helpers readparam (qstem-libs/readparams.cpp:173-218), readCFGCellParams /
// readNextCFGAtom (qstem-libs/fileio_fftw3.cpp:721-776, 908-975), readUnitCell /
// replicateUnitCell (:1313-1657, 1188-1305), rotateVect (qstem-libs/matrixlib.cpp:599-635) and
// getZNumber (qstem-libs/fileio_fftw3.cpp:2299-2321).  Only the keys that reach params_t are
// evaluated (src/rwQsc.cu:937-1001); the rest of the QSTEM vocabulary is parsed by the reference
// into a MULS struct that FDES never looks at again.
//
// Kept from the reference, because they change the numbers:
//   * readparam searches from the CURRENT file position, wraps once, matches the key anywhere in the
//     line (strstr) after cutting the line at '%' -- so "slices:" also matches "center slices:" and
//     the result depends on the order of the lines;
//   * MULS fields and atom coordinates are float32, intermediate arithmetic is double;
//   * dn = n/2 (m = 2n), sub-slice thickness = slice thickness / 10, defocus [nm] -> C1, Cs [mm] ->
//     C3, the astigmatism ANGLE lands in A1_1 scaled by 1e-9 and C5 is scaled by 1e7 * 1e-3
//     (src/rwQsc.cu:965-969); illumination angle = alpha / 1e3; mtf_d and illumination_angle are
//     not read; specimen tilt offsets = crystal tilts (applied a second time by the FDES driver);
//   * atoms are centred by subtracting half the extent (max - min)/2 with max starting at 0 and min
//     at 1 (src/rwQsc.cu:1032-1076).
// Rejected loudly (the reference draws from ran1 / Einstein displacements there, qstem-libs
// fileio_fftw3.cpp:1226-1268, or reads other formats): partial occupancies or shared sites in the
// unit cell, `tds: yes`, `Cube:` mode, .cssr/.dat/.pdb/.xyz specimen files, `mode:` values that do not
// contain "TEM" (the reference tests with strstr, so "STEM" passes as TEM there and here, src/rwQsc.cu:35).
#include "params.h"
#include "emd.h"
#include "qsc.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace fdes {
namespace {

// ---- readparams.cpp: one open parameter file with a position ---------------------------------
struct ParFile {
    std::vector<std::string> lines;   // as fgets(buf, 1024, fp) would return them
    size_t pos = 0;

    bool open(const std::string& path)
    {
        FILE* f = fopen(path.c_str(), "r");
        if (!f) return false;
        char buf[1024];
        lines.clear();
        while (fgets(buf, sizeof buf, f)) lines.emplace_back(buf);
        fclose(f);
        pos = 0;
        return true;
    }
    // fgets at the current position
    bool next_raw(std::string& out)
    {
        if (pos >= lines.size()) return false;
        out = lines[pos++];
        return true;
    }
    // readparam (readparams.cpp:173-218): rest of the first line at or after the current position
    // that contains `title` once cut at '%'; wraps to the start of the file once.
    bool readparam(const char* title, std::string& out, bool wrap = true)
    {
        auto scan = [&]() -> bool {
            while (pos < lines.size()) {
                std::string l = lines[pos++];
                const size_t c = l.find('%');
                if (c != std::string::npos) l.resize(c);
                const size_t t = l.find(title);
                if (t != std::string::npos) { out = l.substr(t + strlen(title)); return true; }
            }
            return false;
        };
        if (scan()) return true;
        if (wrap) { pos = 0; if (scan()) return true; }
        return false;
    }
};

// strnext (readparams.cpp:232-247): next word after a run of delimiters, NULL at end / newline
const char* strnext(const char* str, const char* delim)
{
    bool found = false;
    const char* s = str;
    for (; *s != '\0'; s++) {
        if (strchr(delim, *s)) found = true;
        if (found && strchr(delim, *s) == nullptr) break;
    }
    if (*s == '\0' || *s == '\n') return nullptr;
    return s;
}

bool scan_f(const std::string& s, float& v) { return sscanf(s.c_str(), "%g", &v) == 1; }
bool scan_i(const std::string& s, int& v) { return sscanf(s.c_str(), "%d", &v) == 1; }
bool answer_is(const std::string& s, char c)
{
    char a[256] = "";
    sscanf(s.c_str(), "%255s", a);
    return tolower((unsigned char)a[0]) == c;
}

const char* const kElTable =
    "H HeLiBeB C N O F NeNaMgAlSiP S Cl"
    "ArK CaScTiV CrMnFeCoNiCuZnGaGeAsSeBr"
    "KrRbSrY ZrNbMoTcRuRhPdAgCdInSnSbTe"
    "I XeCsBaLaCePrNdPmSmEuGdTbDyHoErTm"
    "YbLuHfTaW ReOsIrPtAuHgTlPbBiPoAtRn"
    "FrRaAcThPaU NpPuAmCmBkCfEsFmMdNoLr";

// getZNumber (fileio_fftw3.cpp:2299-2321): two-character symbol looked up by strstr in the packed table
int z_number(const std::string& line)
{
    char el[3] = {0, 0, 0};
    if (!line.empty()) el[0] = line[0];
    if (line.size() > 1) el[1] = line[1];
    if (atoi(el + 1) != 0 || el[1] == '\n' || el[1] == '\0' || el[1] == '\r') el[1] = ' ';
    const char* hit = strstr(kElTable, el);
    return hit ? (int)(hit - kElTable) / 2 + 1 : 0;
}

struct CellAtom { float z, y, x, dw, occ; int Z; };

// rotateVect (matrixlib.cpp:599-635)
void rotate_vect(double* u, double px, double py, double pz)
{
    const double M[3][3] = {
        {cos(pz) * cos(py), cos(pz) * sin(py) * sin(px) - sin(pz) * cos(px), cos(pz) * sin(py) * cos(px) + sin(pz) * sin(px)},
        {sin(pz) * cos(py), sin(pz) * sin(py) * sin(px) + cos(pz) * cos(px), sin(pz) * sin(py) * cos(px) - cos(pz) * sin(px)},
        {-sin(py), cos(py) * sin(px), cos(py) * cos(px)}};
    const double o0 = M[0][0] * u[0] + M[0][1] * u[1] + M[0][2] * u[2];
    const double o1 = M[1][0] * u[0] + M[1][1] * u[1] + M[1][2] * u[2];
    const double o2 = M[2][0] * u[0] + M[2][1] * u[1] + M[2][2] * u[2];
    u[0] = o0; u[1] = o1; u[2] = o2;
}

// wavelength (src/rwQsc.cu:1236-1248), Angstrom
double wavelength_A(double kev)
{
    const double emass = 510.99906, hc = 12.3984244;
    return hc / sqrt(kev * (2 * emass + kev));
}

std::string dir_of(const std::string& path)
{
    const size_t s = path.find_last_of('/');
    return s == std::string::npos ? std::string() : path.substr(0, s + 1);
}

// readUnitCell in NCELL mode for a .cfg file (fileio_fftw3.cpp:1313-1657); coordinates in Angstrom.
// ax_by_c receives the size of the (tilted) super cell.
std::vector<CellAtom> read_unit_cell(const std::string& cfg, int ncx, int ncy, int ncz, float ctx, float cty,
                                     float ctz, float xoff, float yoff, float ax_by_c[3])
{
    // readCFGCellParams (:721-776)
    ParFile f;
    if (!f.open(cfg)) throw std::runtime_error("Could not open CFG input file " + cfg);
    std::string buf;
    int ncoord = 0;
    double scale = 0, Mm[3][3] = {{0}};
    if (f.readparam("Number of particles =", buf)) sscanf(buf.c_str(), "%d", &ncoord);
    if (f.readparam("A =", buf)) sscanf(buf.c_str(), "%lf", &scale);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            char key[32];
            snprintf(key, sizeof key, "H0(%d,%d) =", r + 1, c + 1);
            if (f.readparam(key, buf)) sscanf(buf.c_str(), "%lf", &Mm[r][c]);
        }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Mm[r][c] *= scale;
    if (ncoord < 1) throw std::runtime_error("Error reading configuration file " + cfg + " - ncoord =0");
    if (ncx < 1 || ncy < 1 || ncz < 1) throw std::runtime_error(".qsc: NCELLX/Y/Z must be >= 1");

    // readNextCFGAtom (:908-975), atoms stored from the back
    f.pos = 0;
    int entryCount = 3;
    const bool noVelocity = f.readparam(".NO_VELOCITY.", buf);
    if (f.readparam("entry_count =", buf)) sscanf(buf.c_str(), "%d", &entryCount);
    if (!noVelocity) entryCount += 3;
    const int off = 3 * (noVelocity ? 0 : 1);
    std::vector<CellAtom> atoms((size_t)ncoord * ncx * ncy * ncz);
    std::vector<double> data((size_t)entryCount + 1, 0.0);
    double mass = 28;
    int element = 1;
    for (int i = ncoord - 1; i >= 0; i--) {
        std::string line;
        auto need = [&]() { if (!f.next_raw(line)) throw std::runtime_error("number of atoms does not agree with atoms in file " + cfg); };
        need();
        const char* nx = strnext(line.c_str(), " \t");
        if (atof(line.c_str()) >= 1.0 && (nx == nullptr || *nx == '#')) {
            mass = atof(line.c_str());
            need();
            element = z_number(line);
            need();
        }
        const char* s = line.c_str();
        while (*s && strchr(" \t", *s) != nullptr) s++;
        for (int j = 0; j < entryCount; j++) {
            if (s == nullptr) throw std::runtime_error("readNextCFGatom: incomplete data line in " + cfg + ": " + line);
            data[j] = atof(s);
            s = strnext(s, " \t");
        }
        CellAtom a;
        a.Z = element;
        a.x = (float)data[0]; a.y = (float)data[1]; a.z = (float)data[2];
        a.dw = (float)(0.45 * 28.0 / mass);
        a.occ = 1.0f;
        if (entryCount > 3 + off) a.dw = (float)data[3 + off];
        if (entryCount > 4 + off) a.occ = (float)data[4 + off];
        if (a.Z < 1 || a.Z > 103) throw std::runtime_error("bad atomic number in file " + cfg);
        atoms[i] = a;
    }
    // qsort(atoms, ncoord, atomCompareZYX) (:1491-1493, :109-123)
    std::stable_sort(atoms.begin(), atoms.begin() + ncoord, [](const CellAtom& a, const CellAtom& b) {
        if (a.z != b.z) return a.z < b.z;
        if (a.y != b.y) return a.y < b.y;
        return a.x < b.x;
    });
    // replicateUnitCell (:1188-1305) without the vacancy / shared-site lottery
    for (int i = ncoord - 1; i >= 0; i--) {
        if (atoms[i].occ < 1.f)
            throw std::runtime_error(".cfg: partial occupancies are not supported (the reference draws them with ran1)");
        if (i > 0 && fabs(atoms[i].x - atoms[i - 1].x) < 1e-6 && fabs(atoms[i].y - atoms[i - 1].y) < 1e-6 &&
            fabs(atoms[i].z - atoms[i - 1].z) < 1e-6)
            throw std::runtime_error(".cfg: several atoms on one site are not supported (the reference draws one with ran1)");
        for (int icx = ncx - 1; icx >= 0; icx--)
            for (int icy = ncy - 1; icy >= 0; icy--)
                for (int icz = ncz - 1; icz >= 0; icz--) {
                    CellAtom& d = atoms[(size_t)(icz + icy * ncz + icx * ncy * ncz) * ncoord + i];
                    const CellAtom s = atoms[i];
                    d.dw = s.dw; d.occ = s.occ; d.Z = s.Z;
                    d.x = (float)((double)(s.x + (float)icx) + 0.0);
                    d.y = (float)((double)(s.y + (float)icy) + 0.0);
                    d.z = (float)((double)(s.z + (float)icz) + 0.0);
                }
    }
    // fractional -> cartesian (:1523-1538)
    for (CellAtom& a : atoms) {
        const double x = Mm[0][0] * a.x + Mm[1][0] * a.y + Mm[2][0] * a.z;
        const double y = Mm[0][1] * a.x + Mm[1][1] * a.y + Mm[2][1] * a.z;
        const double z = Mm[0][2] * a.x + Mm[1][2] * a.y + Mm[2][2] * a.z;
        a.x = (float)x; a.y = (float)y; a.z = (float)z;
    }
    // box of the rotated super cell (:1560-1590)
    const double bc[3] = {ncx / 2.0, ncy / 2.0, ncz / 2.0};
    double ctr[3];
    for (int c = 0; c < 3; c++) ctr[c] = Mm[0][c] * bc[0] + Mm[1][c] * bc[1] + Mm[2][c] * bc[2];
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    bool first = true;
    for (int icx = 0; icx <= ncx; icx += ncx)
        for (int icy = 0; icy <= ncy; icy += ncy)
            for (int icz = 0; icz <= ncz; icz += ncz) {
                double u[3];
                for (int c = 0; c < 3; c++) u[c] = Mm[0][c] * (icx - bc[0]) + Mm[1][c] * (icy - bc[1]) + Mm[2][c] * (icz - bc[2]);
                rotate_vect(u, ctx, cty, ctz);
                for (int c = 0; c < 3; c++) {
                    const double v = u[c] + ctr[c];
                    if (first) lo[c] = hi[c] = v;
                    else { lo[c] = lo[c] > v ? v : lo[c]; hi[c] = hi[c] < v ? v : hi[c]; }
                }
                first = false;
            }
    if (ctx != 0 || cty != 0 || ctz != 0)
        for (CellAtom& a : atoms) {
            double u[3] = {a.x - ctr[0], a.y - ctr[1], a.z - ctr[2]};
            rotate_vect(u, ctx, cty, ctz);
            a.x = (float)(u[0] + ctr[0]); a.y = (float)(u[1] + ctr[1]); a.z = (float)(u[2] + ctr[2]);
        }
    for (CellAtom& a : atoms) {
        a.x = (float)(a.x - lo[0]); a.y = (float)(a.y - lo[1]); a.z = (float)(a.z - lo[2]);
    }
    for (int c = 0; c < 3; c++) ax_by_c[c] = (float)(hi[c] - lo[c]);
    if (xoff != 0 || yoff != 0)
        for (CellAtom& a : atoms) { a.x += xoff; a.y += yoff; }
    return atoms;
}

}  // namespace

bool is_qsc_name(const char* file) { return file && strstr(file, ".qsc") != nullptr; }

namespace {
bool read_qsc_impl(const char* file, Params& p, Atoms* atoms_out, bool atoms_from_external, float shift[3]);
}

bool read_qsc(const char* file, Params& p, Atoms* atoms_out, bool atoms_from_external)
{
    float shift[3];
    return read_qsc_impl(file, p, atoms_out, atoms_from_external, shift);
}

// Scan raster and detectors of a `mode: STEM` parameter file: the keys readQsc parses into MULS
// (scan_x_start / stop / pixels, scan_y_*: src/rwQsc.cu:444-466; `detector: inner outer name shiftx
// shifty` in mrad: :698-735) and FDES then never uses.  Positions follow QSTEM's raster
// x_i = start + i (stop - start) / pixels in the super-cell frame [A]; they are returned in the frame
// of the atoms after readQsc's centring (src/rwQsc.cu:1071-1076), in metres, x index slowest.
bool read_qsc_scan(const char* file, QscScan& out)
{
    Params p;
    Atoms atoms;
    float shift[3];
    if (!read_qsc_impl(file, p, &atoms, false, shift)) return false;
    ParFile q;
    if (!q.open(file)) return false;
    std::string buf;
    float xs = 0, xe = 0, ys = 0, ye = 0;
    int nx = 0, ny = 0;
    auto needf = [&](const char* key, float& v) {
        if (!q.readparam(key, buf) || !scan_f(buf, v)) throw std::runtime_error(std::string(".qsc: STEM scan needs `") + key + "`");
    };
    auto needi = [&](const char* key, int& v) {
        if (!q.readparam(key, buf) || !scan_i(buf, v)) throw std::runtime_error(std::string(".qsc: STEM scan needs `") + key + "`");
    };
    needf("scan_x_start:", xs); needf("scan_x_stop:", xe); needi("scan_x_pixels:", nx);
    needf("scan_y_start:", ys); needf("scan_y_stop:", ye); needi("scan_y_pixels:", ny);
    if (nx < 1) nx = 1;
    if (ny < 1) ny = 1;
    out.nx = nx; out.ny = ny;
    out.xy.resize(2 * (size_t)nx * ny);
    const float dx = (xe - xs) / (float)nx, dy = (ye - ys) / (float)ny;
    for (int ix = 0; ix < nx; ix++)
        for (int iy = 0; iy < ny; iy++) {
            const float x = xs + (float)ix * dx, y = ys + (float)iy * dy;
            out.xy[2 * ((size_t)ix * ny + iy) + 0] = (float)(x * 1e-10) - shift[0];
            out.xy[2 * ((size_t)ix * ny + iy) + 1] = (float)(y * 1e-10) - shift[1];
        }
    out.det_mrad.clear();
    q.pos = 0;
    while (q.readparam("detector:", buf, false)) {
        float in = 0, outer = 0;
        if (sscanf(buf.c_str(), "%g %g", &in, &outer) == 2) { out.det_mrad.push_back(in); out.det_mrad.push_back(outer); }
    }
    return true;
}

namespace {
bool read_qsc_impl(const char* file, Params& p, Atoms* atoms_out, bool atoms_from_external, float shift[3])
{
    shift[0] = shift[1] = shift[2] = 0.f;
    ParFile q;
    if (!q.open(file)) {
        fprintf(stderr, "could not open input file %s!\n", file);
        return false;
    }
    const double pi = 3.1415926535897;
    std::string buf;
    // mode (src/rwQsc.cu:31-44): STEM when absent -- and STEM is refused like everything but TEM
    bool tem = false;
    if (q.readparam("mode:", buf)) tem = buf.find("TEM") != std::string::npos;
    if (!tem) throw std::runtime_error(".qsc: FDES supports only `mode: TEM` parameter files");
    q.readparam("print level:", buf);
    q.readparam("save level:", buf);
    if (!q.readparam("filename:", buf)) throw std::runtime_error(".qsc: no `filename:` (crystal .cfg file)");
    char tok[1024] = "";
    sscanf(buf.c_str(), "%1023s", tok);
    std::string fileBase(tok);
    if (!fileBase.empty() && fileBase[0] == '"') {
        const size_t a = buf.find('"');
        fileBase = buf.substr(a + 1);
        const size_t b = fileBase.find('"');
        if (b != std::string::npos) fileBase.resize(b);
    }
    q.readparam("wavename:", buf);   // moves the file position like the reference (:66-69)
    int ncx = 0, ncy = 0, ncz = 0, cellDiv = 1;
    if (q.readparam("NCELLX:", buf)) scan_i(buf, ncx);
    if (q.readparam("NCELLY:", buf)) scan_i(buf, ncy);
    if (q.readparam("NCELLZ:", buf)) {
        char a[256] = "";
        sscanf(buf.c_str(), "%255s", a);
        if (char* sl = strchr(a, '/')) { *sl = 0; cellDiv = atoi(sl + 1); }
        ncz = atoi(a);
    }
    auto angle = [&](const char* key) -> float {   // "%g %s", degrees when the unit starts with 'd'
        float v = 0.f;
        if (q.readparam(key, buf)) {
            char unit[256] = "";
            sscanf(buf.c_str(), "%g %255s", &v, unit);
            if (tolower((unsigned char)unit[0]) == 'd') v = (float)(v * (pi / 180.0));
        }
        return v;
    };
    const float btiltx = angle("Beam tilt X:"), btilty = angle("Beam tilt Y:");
    q.readparam("Tilt back:", buf);
    const float ctiltx = angle("Crystal tilt X:"), ctilty = angle("Crystal tilt Y:"), ctiltz = angle("Crystal tilt Z:");
    float cube[3] = {0, 0, 0};
    if (q.readparam("Cube:", buf)) sscanf(buf.c_str(), "%g %g %g", cube, cube + 1, cube + 2);
    if (cube[0] > 0 && cube[1] > 0 && cube[2] > 0)
        throw std::runtime_error(".qsc: `Cube:` specimens (tiltBoxed) are not supported");
    q.readparam("Adjust cube size with tilt:", buf);
    if (q.readparam("tds:", buf) && answer_is(buf, 'y'))
        throw std::runtime_error(".qsc: `tds: yes` (QSTEM Einstein displacements) is not supported; use frozen_phonons:");
    q.readparam("temperature:", buf);
    q.readparam("phonon-File:", buf);
    // specimen file (:169-214): only .cfg; resolved against the cwd like the reference, then
    // against the directory of the .qsc file
    std::string atomPosFile = fileBase;
    if (atomPosFile.find('.') == std::string::npos) atomPosFile += ".cfg";
    if (atomPosFile.size() < 4 || atomPosFile.compare(atomPosFile.size() - 4, 4, ".cfg") != 0)
        throw std::runtime_error(".qsc: only .cfg specimen files are supported (got " + atomPosFile + ")");
    std::string cfgPath = atomPosFile;
    if (FILE* t = fopen(cfgPath.c_str(), "r")) fclose(t);
    else cfgPath = dir_of(file) + atomPosFile;
    float xOffset = 0.f, yOffset = 0.f;
    if (q.readparam("xOffset:", buf)) scan_f(buf, xOffset);
    if (q.readparam("yOffset:", buf)) scan_f(buf, yOffset);
    float cell[3];
    std::vector<CellAtom> cellAtoms = read_unit_cell(cfgPath, ncx, ncy, ncz, ctiltx, ctilty, ctiltz, xOffset, yOffset, cell);
    if (cellAtoms.empty()) throw std::runtime_error(".qsc: no atom within simulation boundaries");

    int nx = 0, ny = 0;
    if (!q.readparam("nx:", buf)) throw std::runtime_error(".qsc: no `nx:`");
    scan_i(buf, nx);
    if (q.readparam("ny:", buf)) scan_i(buf, ny); else ny = nx;
    float resX = 0.f, resY = 0.f, v0 = 0.f;
    if (q.readparam("resolutionX:", buf)) scan_f(buf, resX);
    if (q.readparam("resolutionY:", buf)) scan_f(buf, resY);
    if (!q.readparam("v0:", buf)) throw std::runtime_error(".qsc: no `v0:`");
    scan_f(buf, v0);
    int centerSlices = 0;
    if (q.readparam("center slices:", buf)) centerSlices = answer_is(buf, 'y');
    // slices and slice thickness (:265-308)
    float sliceThickness = 0.f;
    int slices = 0;
    if (q.readparam("slice-thickness:", buf)) {
        scan_f(buf, sliceThickness);
        if (q.readparam("slices:", buf)) scan_i(buf, slices);
        else slices = (int)(cell[2] / (cellDiv * sliceThickness) + 0.99);
        slices += centerSlices;
    } else if (q.readparam("slices:", buf)) {
        scan_i(buf, slices);
        if (slices == 1 && cellDiv == 1) sliceThickness = cell[2] / cellDiv;
        else sliceThickness = cell[2] / (cellDiv * slices);
    }
    if (slices == 0) throw std::runtime_error(".qsc: Number of slices = 0");
    q.readparam("slices between outputs:", buf);
    q.readparam("zOffset:", buf);
    if (resX == 0.f) resX = (float)(cell[0] / (double)nx);
    if (resY == 0.f) resY = (float)(cell[1] / (double)ny);
    // the yes/no block of :344-399 only moves the file position
    for (const char* key : {"periodicXY:", "periodicZ:", "bandlimit f_trans:", "read potential:", "save potential:",
                            "save projected potential:", "plot V(r)*r:", "one time integration:", "potential3D:",
                            "Runs for averaging:", "Store TDS diffr. patt. series:", "potential progress interval:"})
        q.readparam(key, buf);
    // lens (:480-560, :619-620)
    for (const char* key : {"dE/E:", "dI/I:", "dV/V:", "Cc:"}) q.readparam(key, buf);
    float Cs = 0.f, C5 = 0.f;
    if (!q.readparam("Cs:", buf)) throw std::runtime_error(".qsc: no `Cs:`");
    scan_f(buf, Cs);
    Cs = (float)(Cs * 1.0e7);
    if (q.readparam("C5:", buf)) { scan_f(buf, C5); C5 = (float)(C5 * 1.0e7); }
    float df0 = -(float)sqrt(1.5 * Cs * wavelength_A(v0));
    if (q.readparam("defocus:", buf)) {
        char a[256] = "";
        sscanf(buf.c_str(), "%255s", a);
        if (tolower((unsigned char)a[0]) == 's') df0 = -(float)sqrt(1.5 * Cs * wavelength_A(v0));
        else if (tolower((unsigned char)a[0]) == 'o') df0 = -(float)sqrt(Cs * wavelength_A(v0));
        else { scan_f(buf, df0); df0 = (float)(10.0 * df0); }
    }
    float astigMag = 0.f, astigAngle = 0.f;
    if (q.readparam("astigmatism:", buf)) scan_f(buf, astigMag);
    astigMag = (float)(10.0 * astigMag);
    if (q.readparam("astigmatism angle:", buf)) scan_f(buf, astigAngle);
    astigAngle = (float)(astigAngle * (pi / 180.0));
    float alpha = 0.f;
    if (!q.readparam("alpha:", buf)) throw std::runtime_error(".qsc: no `alpha:`");
    scan_f(buf, alpha);

    // MULS -> params_t (src/rwQsc.cu:937-1001)
    p = Params();
    p.cst_pi = 3.1415927f;
    p.n3 = 1;
    p.n1 = nx; p.n2 = ny;
    p.dn1 = (int)roundf((float)(nx / 2));
    p.dn2 = (int)roundf((float)(ny / 2));
    p.m1 = p.n1 + 2 * p.dn1; p.m2 = p.n2 + 2 * p.dn2;
    p.m3 = slices;
    p.d1 = (float)(resX * 1e-10);
    p.d2 = (float)(resY * 1e-10);
    p.d3 = (float)(sliceThickness * 1e-10);
    p.subSlTh = (float)(sliceThickness * 1e-10 / 10);
    p.tilt_off[0] = ctiltx; p.tilt_off[1] = ctilty; p.tilt_off[2] = ctiltz;
    p.tiltspec.assign(2, 0.f); p.tiltbeam.assign(2, 0.f); p.defoci.assign(1, 0.f);
    p.tiltbeam[0] = btiltx; p.tiltbeam[1] = btilty;
    p.E0 = (float)(v0 * 1e3);
    p.illangle = (float)(alpha / 1e3);
    p.ab0[AB_A1] = (float)(astigMag * 1e-9);
    p.ab1[AB_A1] = (float)(astigAngle * 1e-9);
    p.ab0[AB_C1] = (float)(df0 * 1e-10);
    p.ab0[AB_C3] = (float)(Cs * 1e-10);
    p.ab0[AB_C5] = (float)(C5 * 1e-3);
    const size_t dot = atomPosFile.find(".cfg");
    p.material = atomPosFile.substr(0, dot);
    char cellnum[128];
    snprintf(cellnum, sizeof cellnum, "_CELL_%02d_%02d_%02d", ncx, ncy, ncz);
    p.sample_name = atomPosFile.substr(0, dot) + cellnum;
    if (q.readparam("cal_mode:", buf)) scan_i(buf, p.mode);
    if (q.readparam("focus_spread:", buf)) scan_f(buf, p.defocspread);
    if (q.readparam("objective_aperture:", buf)) scan_f(buf, p.ObjAp);
    if (q.readparam("pixel_dose:", buf)) scan_f(buf, p.pD);
    if (q.readparam("absorptive_potential_factor:", buf)) scan_f(buf, p.imPot);
    if (q.readparam("mtf_a:", buf)) scan_f(buf, p.mtfa);
    if (q.readparam("mtf_b:", buf)) scan_f(buf, p.mtfb);
    if (q.readparam("mtf_c:", buf)) scan_f(buf, p.mtfc);
    if (q.readparam("frozen_phonons:", buf)) scan_i(buf, p.frPh);

    if (!atoms_from_external && atoms_out) {
        // A -> m, DWF A^2 -> m^2, then centring (src/rwQsc.cu:1027-1076)
        Atoms& at = *atoms_out;
        at = Atoms();
        const size_t n = cellAtoms.size();
        at.Z.resize(n); at.xyz.resize(3 * n); at.dwf.resize(n); at.occ.resize(n);
        float mn[3] = {1, 1, 1}, mx[3] = {0, 0, 0};
        for (size_t i = 0; i < n; i++) {
            at.Z[i] = cellAtoms[i].Z;
            at.dwf[i] = (float)(cellAtoms[i].dw * 1e-20);
            at.occ[i] = cellAtoms[i].occ;
            const float c[3] = {(float)(cellAtoms[i].x * 1e-10), (float)(cellAtoms[i].y * 1e-10), (float)(cellAtoms[i].z * 1e-10)};
            for (int k = 0; k < 3; k++) {
                at.xyz[3 * i + k] = c[k];
                if (c[k] > mx[k]) mx[k] = c[k];
                if (c[k] < mn[k]) mn[k] = c[k];
            }
        }
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) at.xyz[3 * i + k] = at.xyz[3 * i + k] - (mx[k] - mn[k]) / 2;
        for (int k = 0; k < 3; k++) shift[k] = (mx[k] - mn[k]) / 2;
        p.nAt = (int)n;
    }
    consistent_params(p);
    return true;
}
}  // namespace

bool read_input(const char* file, Params& p, Atoms* atoms, bool atoms_from_external)
{
    // same order of tests as src/FDESExport.cu:85-102: .emd, .cnf, .qsc
    if (file && strstr(file, ".emd")) return read_emd(file, p, atoms, atoms_from_external);
    if (is_qsc_name(file) && !(strstr(file, ".cnf"))) return read_qsc(file, p, atoms, atoms_from_external);
    return read_cnf(file, p, atoms, atoms_from_external);
}

}  // namespace fdes
