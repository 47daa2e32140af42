// TEST INFRASTRUCTURE ONLY (oracle build): stand-in for ompl::base::ReedsSheppStateSpace exposing
// distance() only (rs_statespace.h:23,34,52). The arithmetic is oracle/port/ompl_rs_distance.h.
#pragma once
#include "ompl/base/spaces/SE2StateSpace.h"
#include "../../../../port/ompl_rs_distance.h"
namespace ompl
{
    namespace base
    {
        class ReedsSheppStateSpace : public SE2StateSpace
        {
        public:
            explicit ReedsSheppStateSpace(double turningRadius = 1.0) : rho_(turningRadius) {}
            double distance(const State *s1, const State *s2) const
            {
                const StateType *a = static_cast<const StateType *>(s1), *b = static_cast<const StateType *>(s2);
                return ompl_rs::distance(rho_, a->getX(), a->getY(), a->getYaw(), b->getX(), b->getY(), b->getYaw());
            }

        protected:
            double rho_;
        };
    }
}
