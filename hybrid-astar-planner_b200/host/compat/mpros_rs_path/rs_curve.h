// Drop-in for mpros_rs_path/include/mpros_rs_path/rs_curve.h: namespace RS with RSCurve, Path, PathSeg, Point and
// PathFunction::path1..12 backed by the CUDA library (see compat/mpros_navigation_map/nav_map.h).
#ifndef RS_CURVE_H
#define RS_CURVE_H
#include "mpros_facade.h"
namespace RS = mpros_b200::RS;
#endif
