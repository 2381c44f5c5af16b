"""Host driver: the reference's time loop around ``s_tvd_rk``.

Restates ``p_main`` / ``s_perform_time_step`` (reference ``src/simulation/p_main.fpp:45-91``,
``src/simulation/m_start_up.fpp:566-642``): fixed ``dt`` with the last step clipped to
``finaltime = t_step_stop*dt``; the state object only needs ``step(dt)``.
"""


def time_steps(case):
    """Yield the dt of every step from t_step_start to t_step_stop exactly as the reference does."""
    dt = float(case["dt"])
    t0 = int(case.get("t_step_start", 0))
    t1 = int(case["t_step_stop"])
    mytime = 0.0 if t0 == 0 else t0 * dt
    finaltime = t1 * dt
    for _ in range(t0, t1):
        if (mytime + dt) >= finaltime:
            dt = finaltime - mytime
        yield dt
        mytime = mytime + dt


def run(state, case, on_step=None):
    for it, dt in enumerate(time_steps(case)):
        state.step(dt)
        if on_step is not None:
            on_step(it, dt)
    return state
