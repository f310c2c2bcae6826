/* TEST INFRASTRUCTURE (oracle/_ref build only).
 * Boost is absent from this image; qstem-libs only needs boost::shared_ptr
 * (qstem-libs/data_containers.h:73,99), which std::shared_ptr satisfies. */
#ifndef FDES_B200_ORACLE_BOOST_SHIM_H
#define FDES_B200_ORACLE_BOOST_SHIM_H
#include <memory>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
#endif
