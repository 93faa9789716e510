"""Simulator options: the subset of ahkab/options.py:38-209 the batched
Newton/MNA path reads.  Same names and defaults; `snapshot()` freezes them
into the POD `ahk_options` that crosses the C-ABI (include/ahkab_b200.h)."""

# tolerances (ahkab/options.py:50-56)
vea = 1e-6
ver = 1e-3
iea = 1e-9
ier = 1e-3
# minimum conductance to ground (options.py:60)
gmin = 1e-12
# Newton damping (options.py:71-80)
nr_damp_first_iters = False
nl_voltages_lock = True
nl_voltages_lock_factor = 4
# solve methods (options.py:83-89)
use_standard_solve_method = True
use_gmin_stepping = True
use_source_stepping = True
# DC (options.py:99-112)
dc_max_nr_iter = 10000
dc_use_guess = True
dc_log_step = 'LOG'
dc_lin_step = 'LIN'
dc_sweep_skip_allowed = True
# transient (options.py:115-132)
default_tran_method = "TRAP"
hmin = 1e-20
transient_max_time_iter = 0
transient_max_nr_iter = 20
transient_prediction_as_x0 = True
transient_use_aposteriori_step_control = True
transient_aposteriori_step_threshold = 0.9
transient_no_step_control = False
cmin = 1e-18
# AC (options.py:160-166)
ac_log_step = 'LOG'
ac_lin_step = 'LIN'
ac_max_nr_iter = 20
ac_phase_in_deg = False

_FIELDS = ('vea', 'ver', 'iea', 'ier', 'gmin', 'cmin', 'hmin',
           'nl_voltages_lock_factor', 'transient_aposteriori_step_threshold',
           'nl_voltages_lock', 'nr_damp_first_iters', 'dc_max_nr_iter',
           'transient_max_nr_iter', 'ac_max_nr_iter',
           'use_standard_solve_method', 'use_gmin_stepping',
           'use_source_stepping', 'transient_prediction_as_x0',
           'transient_use_aposteriori_step_control', 'transient_max_time_iter')


def snapshot():
    """Current option values as a plain dict (keys = ahk_options fields)."""
    g = globals()
    return {k: g[k] for k in _FIELDS}
