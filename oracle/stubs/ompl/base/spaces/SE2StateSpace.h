// TEST INFRASTRUCTURE ONLY (oracle build): the sliver of ompl::base::SE2StateSpace that
// mpros_utils/rs_statespace.h:37-48 touches (allocState, setXY, setYaw).
#pragma once
namespace ompl
{
    namespace base
    {
        struct State
        {
        };
        class SE2StateSpace
        {
        public:
            struct StateType : public State
            {
                double x_ = 0, y_ = 0, yaw_ = 0;
                void setXY(double x, double y)
                {
                    x_ = x;
                    y_ = y;
                }
                void setYaw(double yaw) { yaw_ = yaw; }
                double getX() const { return x_; }
                double getY() const { return y_; }
                double getYaw() const { return yaw_; }
            };
            virtual ~SE2StateSpace() {}
            State *allocState() const { return new StateType(); }
        };
    }
}
