// TEST INFRASTRUCTURE (oracle/_ref build only).
//
// Restatement of the slice of Boost.Odeint the reference uses (Boost is neither vendored in the
// reference nor installed in this image; pinned versions: >=1.57 CMakeLists.txt:90, 1.74 Dockerfile:1-4,
// 1.84.0 Dockerfile.python:17):
//   * runge_kutta_cash_karp54<double>           (cpp/src/spectrum.cpp:27, cpp/src/arguments.cpp:71,109,141,167,215,322)
//   * make_controlled(eps_abs, eps_rel, stepper) (cpp/src/spectrum.cpp:30)
//   * integrate_adaptive(stepper, f, x, t0, t1, dt), found by ADL, for a controlled and for a plain stepper
// State is a scalar (vector_space_algebra): every sum is evaluated left to right with the Butcher
// coefficient pre-multiplied by dt, exactly as explicit_error_generic_rk / generic_rk_scale_sum do.
// Written from knowledge of the published Boost sources; it cannot be diffed against Boost offline,
// so last-bit parity with a real Boost build is "unpinned" (oracle/README.md).
#pragma once
#include <cmath>
#include <cstddef>
#include <limits>
#include <stdexcept>

namespace boost { namespace numeric { namespace odeint {

enum controlled_step_result { success, fail };

template <class State, class Value = double, class Deriv = State, class Time = Value>
class runge_kutta_cash_karp54 {
public:
    typedef State state_type;
    typedef Value value_type;
    typedef Deriv deriv_type;
    typedef Time time_type;
    static const unsigned short order_value = 5;
    static const unsigned short stepper_order_value = 5;
    static const unsigned short error_order_value = 4;
    unsigned short order() const { return order_value; }
    unsigned short stepper_order() const { return stepper_order_value; }
    unsigned short error_order() const { return error_order_value; }

private:
    static Value q(int num, int den) { return static_cast<Value>(num) / static_cast<Value>(den); }
    Value a1[1], a2[2], a3[3], a4[4], a5[5], b[6], db[6], c[6];

    // One Cash-Karp step. x_out may alias in. xerr may be null.
    template <class System>
    void step(System& sys, const State& in, const Deriv& dxdt, Time t, State& out, Time dt, State* xerr) {
        Deriv F[5];
        State x_tmp;
        // stage 1 (no evaluation: dxdt is given), builds x_tmp for stage 2
        x_tmp = 1 * in + (a1[0] * dt) * dxdt;
        sys(x_tmp, F[0], t + c[1] * dt);
        x_tmp = 1 * in + (a2[0] * dt) * dxdt + (a2[1] * dt) * F[0];
        sys(x_tmp, F[1], t + c[2] * dt);
        x_tmp = 1 * in + (a3[0] * dt) * dxdt + (a3[1] * dt) * F[0] + (a3[2] * dt) * F[1];
        sys(x_tmp, F[2], t + c[3] * dt);
        x_tmp = 1 * in + (a4[0] * dt) * dxdt + (a4[1] * dt) * F[0] + (a4[2] * dt) * F[1] + (a4[3] * dt) * F[2];
        sys(x_tmp, F[3], t + c[4] * dt);
        x_tmp = 1 * in + (a5[0] * dt) * dxdt + (a5[1] * dt) * F[0] + (a5[2] * dt) * F[1] + (a5[3] * dt) * F[2] +
                (a5[4] * dt) * F[3];
        sys(x_tmp, F[4], t + c[5] * dt);
        const State in_copy = in;
        out = 1 * in_copy + (b[0] * dt) * dxdt + (b[1] * dt) * F[0] + (b[2] * dt) * F[1] + (b[3] * dt) * F[2] +
              (b[4] * dt) * F[3] + (b[5] * dt) * F[4];
        if (xerr) {
            *xerr = (db[0] * dt) * dxdt + (db[1] * dt) * F[0] + (db[2] * dt) * F[1] + (db[3] * dt) * F[2] +
                    (db[4] * dt) * F[3] + (db[5] * dt) * F[4];
        }
    }

public:
    runge_kutta_cash_karp54() {
        a1[0] = q(1, 5);
        a2[0] = q(3, 40); a2[1] = q(9, 40);
        a3[0] = q(3, 10); a3[1] = q(-9, 10); a3[2] = q(6, 5);
        a4[0] = q(-11, 54); a4[1] = q(5, 2); a4[2] = q(-70, 27); a4[3] = q(35, 27);
        a5[0] = q(1631, 55296); a5[1] = q(175, 512); a5[2] = q(575, 13824); a5[3] = q(44275, 110592); a5[4] = q(253, 4096);
        b[0] = q(37, 378); b[1] = static_cast<Value>(0); b[2] = q(250, 621); b[3] = q(125, 594);
        b[4] = static_cast<Value>(0); b[5] = q(512, 1771);
        db[0] = q(37, 378) - q(2825, 27648);
        db[1] = static_cast<Value>(0);
        db[2] = q(250, 621) - q(18575, 48384);
        db[3] = q(125, 594) - q(13525, 55296);
        db[4] = q(-277, 14336);
        db[5] = q(512, 1771) - q(1, 4);
        c[0] = static_cast<Value>(0); c[1] = q(1, 5); c[2] = q(3, 10); c[3] = q(3, 5);
        c[4] = static_cast<Value>(1); c[5] = q(7, 8);
    }

    // explicit_error_stepper_base::do_step(system, x, t, dt): used when the error stepper is passed as a plain stepper
    template <class System>
    void do_step(System system, State& x, Time t, Time dt) {
        Deriv dxdt;
        system(x, dxdt, t);
        step(system, x, dxdt, t, x, dt, static_cast<State*>(nullptr));
    }
    // do_step(system, in, dxdt, t, out, dt, xerr)
    template <class System>
    void do_step(System system, const State& in, const Deriv& dxdt, Time t, State& out, Time dt, State& xerr) {
        step(system, in, dxdt, t, out, dt, &xerr);
    }
};

template <class Stepper>
class controlled_runge_kutta {
public:
    typedef typename Stepper::state_type state_type;
    typedef typename Stepper::value_type value_type;
    typedef typename Stepper::deriv_type deriv_type;
    typedef typename Stepper::time_type time_type;

private:
    Stepper m_stepper;
    value_type m_eps_abs, m_eps_rel, m_a_x, m_a_dxdt;
    time_type m_max_dt;

    // default_error_checker::error with vector_space_algebra on a scalar state
    value_type error(const state_type& x_old, const deriv_type& dxdt_old, state_type& x_err, time_type dt) const {
        using std::abs;
        const value_type a_dxdt_dt = m_a_dxdt * abs(dt);
        x_err = abs(x_err) / (m_eps_abs + m_eps_rel * (m_a_x * abs(x_old) + a_dxdt_dt * abs(dxdt_old)));
        return abs(x_err);
    }
    // default_step_adjuster
    time_type decrease_step(time_type dt, const value_type error, const int error_order) const {
        using std::pow;
        const value_type f = static_cast<value_type>(9) / static_cast<value_type>(10) *
                             pow(error, static_cast<value_type>(-1) / (error_order - 1));
        const value_type lim = static_cast<value_type>(1) / static_cast<value_type>(5);
        dt *= (f < lim) ? lim : f;  // std::max(f, lim): returns f unless f < lim
        return dt;
    }
    time_type increase_step(time_type dt, value_type error, const int stepper_order) const {
        using std::pow;
        if (error < 0.5) {
            const value_type floor_ = static_cast<value_type>(pow(static_cast<value_type>(5.0), -static_cast<value_type>(stepper_order)));
            error = (floor_ < error) ? error : floor_;  // std::max(floor_, error)
            dt *= static_cast<value_type>(9) / static_cast<value_type>(10) * pow(error, static_cast<value_type>(-1) / stepper_order);
        }
        return dt;
    }

public:
    controlled_runge_kutta(value_type eps_abs, value_type eps_rel, const Stepper& stepper)
        : m_stepper(stepper), m_eps_abs(eps_abs), m_eps_rel(eps_rel), m_a_x(1), m_a_dxdt(1), m_max_dt(0) {}

    template <class System>
    controlled_step_result try_step(System system, state_type& x, time_type& t, time_type& dt) {
        deriv_type dxdt;
        system(x, dxdt, t);
        state_type xnew, xerr;
        m_stepper.do_step(system, x, dxdt, t, xnew, dt, xerr);
        const value_type max_rel_err = error(x, dxdt, xerr, dt);
        if (max_rel_err > static_cast<value_type>(1.0)) {
            dt = decrease_step(dt, max_rel_err, m_stepper.error_order());
            return fail;
        }
        t += dt;
        dt = increase_step(dt, max_rel_err, m_stepper.stepper_order());
        x = xnew;
        return success;
    }
};

template <class Stepper>
controlled_runge_kutta<Stepper> make_controlled(typename Stepper::value_type abs_error,
                                                typename Stepper::value_type rel_error, const Stepper& stepper = Stepper()) {
    return controlled_runge_kutta<Stepper>(abs_error, rel_error, stepper);
}

namespace detail_shim {
template <class T> bool less_with_sign(T t1, T t2, T dt) {
    if (dt > 0) return t2 - t1 > std::numeric_limits<T>::epsilon();
    return t1 - t2 > std::numeric_limits<T>::epsilon();
}
template <class T> bool less_eq_with_sign(T t1, T t2, T dt) {
    if (dt > 0) return t1 - t2 <= std::numeric_limits<T>::epsilon();
    return t2 - t1 <= std::numeric_limits<T>::epsilon();
}
}  // namespace detail_shim

// integrate_adaptive, controlled_stepper_tag
template <class Stepper, class System, class State, class Time>
size_t integrate_adaptive(controlled_runge_kutta<Stepper> stepper, System system, State& start_state, Time start_time,
                          Time end_time, Time dt) {
    size_t count = 0;
    size_t fails = 0;
    while (detail_shim::less_with_sign(start_time, end_time, dt)) {
        if (detail_shim::less_with_sign(end_time, static_cast<Time>(start_time + dt), dt)) {
            dt = end_time - start_time;
        }
        controlled_step_result res;
        do {
            res = stepper.try_step(system, start_state, start_time, dt);
            if (res == fail && ++fails >= 500) {
                throw std::overflow_error("Max number of iterations exceeded (500). A new step size was not found.");
            }
        } while (res == fail);
        fails = 0;
        ++count;
    }
    return count;
}

// integrate_adaptive, stepper_tag: constant-step loop (integrate_const) plus a closing partial step
template <class State, class Value, class Deriv, class TimeS, class System, class Time>
size_t integrate_adaptive(runge_kutta_cash_karp54<State, Value, Deriv, TimeS> stepper, System system, State& start_state,
                          Time start_time, Time end_time, Time dt) {
    Time time = start_time;
    int step = 0;
    while (detail_shim::less_eq_with_sign(static_cast<Time>(time + dt), end_time, dt)) {
        stepper.do_step(system, start_state, time, dt);
        ++step;
        time = start_time + static_cast<Time>(step) * dt;
    }
    size_t steps = static_cast<size_t>(step);
    Time end = start_time + dt * steps;
    if (detail_shim::less_with_sign(end, end_time, dt)) {
        stepper.do_step(system, start_state, end, end_time - end);
        steps++;
    }
    return steps;
}

}}}  // namespace boost::numeric::odeint
