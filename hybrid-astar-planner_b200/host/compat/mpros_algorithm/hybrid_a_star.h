// Drop-in for mpros_algorithm/include/mpros_algorithm/hybrid_a_star.h (see compat/mpros_navigation_map/nav_map.h).
#ifndef HYBRID_A_STAR_H
#define HYBRID_A_STAR_H
#include "mpros_navigation_map/nav_map.h"
#include "mpros_rs_path/rs_curve.h"
class HybridAstar : public mpros_b200::HybridAstar
{
public:
    using mpros_b200::HybridAstar::HybridAstar;
};
#endif
