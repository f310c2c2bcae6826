/* TEST INFRASTRUCTURE (oracle/_ref build only).
 * Minimal stand-in for <hdf5.h>: the reference headers include it
 * (include/crystalMaker.h:47, src/paramStructure.cu:38) but the hot path never
 * calls libhdf5, which is absent from this image.  Only the three handle
 * typedefs the reference's declarations mention are provided. */
#ifndef FDES_B200_ORACLE_HDF5_SHIM_H
#define FDES_B200_ORACLE_HDF5_SHIM_H
typedef long long hid_t;
typedef unsigned long long hsize_t;
typedef int herr_t;
#endif
