// fdes_b200 -- STEM scan description of a QSTEM .qsc file (see qsc.cpp)
#pragma once
#include <vector>

namespace fdes {

struct QscScan {
    int nx = 0, ny = 0;
    std::vector<float> xy;         // [nx][ny][2] probe positions [m] in the frame of the centred atoms
    std::vector<float> det_mrad;   // [ndet][2] inner, outer detector angle [mrad]
};
bool read_qsc_scan(const char* file, QscScan& out);

}  // namespace fdes
