// fdes_b200 -- EMD (HDF5) files without libhdf5: the results / configuration files the reference
// writes with writeHdf5 (src/rwHdf5.cu:27-1084, 1085-1945) and reads with readHdf5 (:1946-2571).
// The byte layout is the one libhdf5 1.8 produces for those calls (superblock version 0, version-1
// object headers, symbol-table groups, contiguous datasets); see emd.cpp.
#pragma once
#include "params.h"

namespace fdes {

// Results file (9-argument writeHdf5).  image [n3][n2][n1]; potential [pot_slices][m2][m1][2] or
// NULL (reference: print level > 0); exitwave [n3][m2][m1][2] or NULL (print level > 1).  Without
// any array this is the configuration file of the 6-argument overload ("config.emd").
// Arrays are stored transposed exactly like the reference: /data/images/data [n1][n2][n3],
// /data/exit_wave/data [m1][m2][n3][2], /data/potential_slices/data [m1][m2][m3][2].
bool write_emd(const char* file, const Params& p, const Atoms& atoms, const float* image, const float* potential,
               int pot_slices, const float* exitwave);

// readHdf5: parameters and (unless atoms_from_external) atoms of an .emd written by the reference
// (libhdf5) or by write_emd.  Throws std::runtime_error on a file it cannot interpret.
bool read_emd(const char* file, Params& p, Atoms* atoms, bool atoms_from_external);

}  // namespace fdes
