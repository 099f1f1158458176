/* libiop_b200.hpp — umbrella header of the C++ adapter (link with libiop_b200/lib/libb200iop.so) */
#ifndef LIBIOP_B200_HPP_
#define LIBIOP_B200_HPP_
#include "b200_device.hpp"
#include "b200_fields.hpp"
#include "b200_field_subset.hpp"
#include "b200_hotpath.hpp"
#include "b200_merkle_tree.hpp"
#endif
