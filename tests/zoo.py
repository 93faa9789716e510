"""Small circuits that exercise every element type / analysis option of the hot path.

Each entry: build(Circuit, tf) -> circuit (tf = the time_functions module of whichever
package builds it: the reference's in oracle/make_golden_misc.py, ours in the tests),
the analysis to run, and the options overrides.  Shared by the golden generator (live
reference) and the parity tests (oracle, host emulation, GPU)."""


def rectifier(Circuit, tf):
    """diode with series resistance (inner Newton, diode.py:359-365) + sin source,
    TRAP with LTE step control."""
    c = Circuit('rectifier')
    c.add_model('diode', 'dr', dict(IS=2e-14, N=1.2, RS=0.5))
    c.add_vsource('V1', 'in', c.gnd, dc_value=0., function=tf.sin(0., 5., 1e3))
    c.add_diode('D1', 'in', 'out', 'dr')
    c.add_resistor('RL', 'out', c.gnd, 1e3)
    c.add_capacitor('CL', 'out', c.gnd, 10e-6)
    return c


def mosq_inverter(Circuit, tf):
    """square-law MOS (stamp path, D/S swap) — DC transfer curve."""
    c = Circuit('mosq inverter')
    c.add_model('mosq', 'nm', dict(TYPE='n', VTO=.5, KP=60e-6, LAMBDA=.03))
    c.add_model('mosq', 'pm', dict(TYPE='p', VTO=-.5, KP=25e-6, LAMBDA=.05))
    c.add_vsource('VDD', 'dd', c.gnd, dc_value=2.5)
    c.add_vsource('VIN', 'in', c.gnd, dc_value=0.)
    c.add_mos('M1', 'out', 'in', 'dd', 'dd', w=4e-6, l=1e-6, model_label='pm')
    c.add_mos('M2', 'out', 'in', c.gnd, c.gnd, w=2e-6, l=1e-6, model_label='nm')
    c.add_resistor('RL', 'out', c.gnd, 1e6)
    return c


def switch_rc(Circuit, tf):
    """voltage-controlled switch with hysteresis driven by a pulse, IE with step control."""
    c = Circuit('switch')
    c.add_model('sw', 'swm', dict(VT=1.0, VH=0.2, RON=10., ROFF=1e6))
    c.add_vsource('VDD', 'dd', c.gnd, dc_value=3.)
    c.add_vsource('VC', 'ctl', c.gnd, dc_value=0.,
                  function=tf.pulse(0., 2., 1e-6, 1e-6, 3e-6, 1e-6, 10e-6))
    c.add_switch('S1', 'dd', 'out', 'ctl', c.gnd, False, 'swm')
    c.add_resistor('RL', 'out', c.gnd, 1e3)
    c.add_capacitor('CL', 'out', c.gnd, 1e-9)
    return c


def controlled(Circuit, tf):
    """E, G, H, F controlled sources + I source — linear OP."""
    c = Circuit('controlled sources')
    c.add_vsource('V1', 'a', c.gnd, dc_value=2.)
    c.add_resistor('R1', 'a', 'b', 1e3)
    c.add_resistor('R2', 'b', c.gnd, 2e3)
    c.add_vcvs('E1', 'c', c.gnd, 'b', c.gnd, 3.)
    c.add_resistor('R3', 'c', 'd', 500.)
    c.add_vccs('G1', 'd', c.gnd, 'a', 'b', 2e-3)
    c.add_resistor('R4', 'd', c.gnd, 1e3)
    c.add_ccvs('H1', 'e', c.gnd, 'V1', 50.)
    c.add_resistor('R5', 'e', 'f', 100.)
    c.add_cccs('F1', 'f', c.gnd, 'V1', 4.)
    c.add_resistor('R6', 'f', c.gnd, 200.)
    c.add_isource('I1', 'b', c.gnd, dc_value=1e-3)
    return c


def coupled(Circuit, tf):
    """coupled inductors (M = K sqrt(L1 L2)) — AC transfer."""
    c = Circuit('transformer')
    c.add_vsource('V1', 'in', c.gnd, dc_value=0., ac_value=1.)
    c.add_resistor('RS', 'in', 'p', 50.)
    c.add_inductor('L1', 'p', c.gnd, 1e-3)
    c.add_inductor('L2', 's', c.gnd, 4e-3)
    c.add_inductor_coupling('K1', 'L1', 'L2', 0.9)
    c.add_resistor('RL', 's', c.gnd, 200.)
    c.add_capacitor('CL', 's', c.gnd, 1e-8)
    return c


def rc_gear3(Circuit, tf):
    """RC low-pass driven by a pulse: GEAR3 with LTE control (tests/tran_gear3 style),
    also run with IE and GEAR5."""
    c = Circuit('rc pulse')
    c.add_vsource('V1', 'in', c.gnd, dc_value=0.,
                  function=tf.pulse(0., 1., 1e-6, 1e-7, 4e-6, 1e-7, 10e-6))
    c.add_resistor('R1', 'in', 'out', 1e3)
    c.add_capacitor('C1', 'out', c.gnd, 1e-9)
    c.add_resistor('R2', 'out', 'o2', 2e3)
    c.add_capacitor('C2', 'o2', c.gnd, 0.5e-9)
    return c


def pwl_rc(Circuit, tf):
    """piece-wise linear source with delay and repetition into an RC network, GEAR2."""
    c = Circuit('pwl')
    x = (0., 1e-6, 1.5e-6, 3e-6, 3.5e-6)
    y = (0., 0., 2., 2., 0.)      # continuous across the repetition wrap (y(3.5u) = y(1u))
    c.add_vsource('V1', 'in', c.gnd, dc_value=0.,
                  function=tf.pwl(x, y, repeat=True, repeat_time=1e-6, td=0.5e-6))
    c.add_resistor('R1', 'in', 'out', 200.)
    c.add_capacitor('C1', 'out', c.gnd, 2e-9)
    c.add_resistor('R2', 'out', c.gnd, 2e3)
    return c


def acnl(Circuit, tf):
    """tests/acnl/test_acnl.py: common-source EKV amplifier — AC linearised around the
    operating point (ac.py:417-443 _generate_J)."""
    c = Circuit('CS amplifier')
    c.add_vsource('vin', '1', '0', dc_value=3, ac_value=1)
    c.add_vsource('vdd', '3', '0', dc_value=30)
    c.add_resistor('Rd', '3', '2', 10e3)
    c.add_capacitor('Cd', '3', '2', 40e-12)
    c.add_resistor('Rs', '4', '0', 1e3)
    c.add_capacitor('Cs', '4', '0', 4e-6)
    c.add_model('ekv', 'ekv0', {'TYPE': 'n', 'VTO': .4, 'KP': 1e-2})
    c.add_mos('m1', '2', '1', '4', '0', w=100e-6, l=1e-6, model_label='ekv0')
    return c


def diode_ac(Circuit, tf):
    """diode + series resistance biased through a resistor: small-signal AC of a diode
    (diode.py:get_drive_ports / g), second non-linear AC case."""
    c = Circuit('diode ac')
    c.add_model('diode', 'dm', dict(IS=1e-14, N=1.5, RS=2.))
    c.add_vsource('V1', 'a', c.gnd, dc_value=1.2, ac_value=1.)
    c.add_resistor('R1', 'a', 'b', 1e3)
    c.add_diode('D1', 'b', 'c', 'dm')
    c.add_resistor('R2', 'c', c.gnd, 100.)
    c.add_capacitor('C1', 'c', c.gnd, 1e-9)
    return c


def am_src(Circuit, tf):
    """tests/amckt: AM voltage / current sources into resistors (time_functions.py:672-728)."""
    c = Circuit('am')
    c.add_vsource('V1', '1', c.gnd, dc_value=0., function=tf.am(sa=10, oc=1, fm=100, fc=1e3, td=1e-3))
    c.add_resistor('R1', '1', c.gnd, 1.)
    c.add_vsource('V2', '2', c.gnd, dc_value=0., function=tf.am(sa=2.5, oc=4, fm=100, fc=1e3, td=1e-3))
    c.add_resistor('R2', '2', c.gnd, 1.)
    c.add_isource('I3', '3', c.gnd, dc_value=0., function=tf.am(sa=10, oc=1, fm=1e3, fc=100, td=1e-3))
    c.add_resistor('R3', '3', c.gnd, 1.)
    c.add_capacitor('C3', '3', c.gnd, 1e-5)
    return c


def sffm_src(Circuit, tf):
    """tests/sffmckt: single-frequency FM sources (time_functions.py:610-670)."""
    c = Circuit('sffm')
    c.add_vsource('V1', '1', c.gnd, dc_value=0., function=tf.sffm(vo=0, va=1e-3, fc=20e3, mdi=10, fs=5e3, td=2e-5))
    c.add_resistor('R1', '1', '1b', 1.)
    c.add_capacitor('C1', '1b', c.gnd, 1e-6)
    c.add_isource('I2', '2', c.gnd, dc_value=0., function=tf.sffm(vo=0, va=1, fc=20e3, mdi=10, fs=5e3, td=0.))
    c.add_resistor('R2', '2', c.gnd, 1.)
    return c


def exp_src(Circuit, tf):
    """tests/time_functions: exponential source (time_functions.py:534-608) into an RC."""
    c = Circuit('exp')
    c.add_vsource('V3', 'in', c.gnd, dc_value=0.,
                  function=tf.exp(v1=.5, v2=-.05, td1=10e-6, tau1=20e-6, td2=400e-6, tau2=30e-6))
    c.add_resistor('R1', 'in', 'out', 1e3)
    c.add_capacitor('C1', 'out', c.gnd, 1e-8)
    return c


def diode_ladder(Circuit, tf, N=100):
    """Non-linear circuit with more than 64 unknowns (the reference solves any size densely,
    dc_analysis.py:796-822): a resistor ladder with a diode and an RC load at every node, some
    diodes bridging back two nodes, driven by a source with a DC value and a sine waveform."""
    c = Circuit('diode ladder')
    c.add_model('diode', 'dl', dict(IS=1e-14, N=1.2))
    c.add_vsource('V1', 'n0', c.gnd, dc_value=5., ac_value=1., function=tf.sin(5., 2., 2e5))
    for i in range(N):
        c.add_resistor('R%d' % i, 'n%d' % i, 'n%d' % (i + 1), 100. + 3 * i)
        c.add_diode('D%d' % i, 'n%d' % (i + 1), c.gnd if i % 3 else 'n%d' % max(0, i - 1), 'dl')
        c.add_resistor('G%d' % i, 'n%d' % (i + 1), c.gnd, 5e3)
        c.add_capacitor('C%d' % i, 'n%d' % (i + 1), c.gnd, 1e-10)
    return c


def diode_mesh(Circuit, tf, N=72):
    """Dense non-linear circuit, n > 64: every pair of the N nodes is joined by a resistor (a
    full MNA matrix), a diode and a resistor from every node to ground, two sources."""
    c = Circuit('diode mesh')
    c.add_model('diode', 'dm', dict(IS=2e-14, N=1.1))
    c.add_vsource('V1', 'm0', c.gnd, dc_value=3.)
    c.add_vsource('V2', 'm1', c.gnd, dc_value=-1.5)
    for i in range(N):
        for j in range(i + 1, N):
            c.add_resistor('R%d_%d' % (i, j), 'm%d' % i, 'm%d' % j, 1e3 * (1 + ((7 * i + 3 * j) % 11)))
        if i >= 2:
            c.add_diode('D%d' % i, 'm%d' % i, c.gnd, 'dm')
            c.add_resistor('G%d' % i, 'm%d' % i, c.gnd, 2e3 + 10 * i)
    return c


ZOO = {
    'rectifier': dict(build=rectifier, an=dict(kind='tran', tstart=0., tstop=2e-3, tstep=1e-5,
                                               method='TRAP', use_step_control=True, x0='op')),
    'mosq_inverter': dict(build=mosq_inverter, an=dict(kind='dc', start=0., stop=2.5, points=26,
                                                       source='VIN')),
    'switch_rc': dict(build=switch_rc, an=dict(kind='tran', tstart=0., tstop=12e-6, tstep=1e-7,
                                               method='IMPLICIT_EULER', use_step_control=True,
                                               x0='op')),
    'controlled': dict(build=controlled, an=dict(kind='op')),
    'coupled': dict(build=coupled, an=dict(kind='ac', start=1e2, stop=1e6, points=41)),
    'rc_gear3': dict(build=rc_gear3, an=dict(kind='tran', tstart=0., tstop=12e-6, tstep=2e-7,
                                             method='GEAR3', use_step_control=True, x0='op')),
    'rc_gear5': dict(build=rc_gear3, an=dict(kind='tran', tstart=0., tstop=12e-6, tstep=2e-7,
                                             method='GEAR5', use_step_control=True, x0='op')),
    'pwl_rc': dict(build=pwl_rc, an=dict(kind='tran', tstart=0., tstop=10e-6, tstep=1e-7,
                                         method='GEAR2', use_step_control=True, x0='op')),
    'rc_gear1': dict(build=rc_gear3, an=dict(kind='tran', tstart=0., tstop=12e-6, tstep=2e-7,
                                             method='GEAR1', use_step_control=True, x0='op')),
    'rc_gear4': dict(build=rc_gear3, an=dict(kind='tran', tstart=0., tstop=12e-6, tstep=2e-7,
                                             method='GEAR4', use_step_control=True, x0='op')),
    'rc_gear6': dict(build=rc_gear3, an=dict(kind='tran', tstart=0., tstop=12e-6, tstep=2e-7,
                                             method='GEAR6', use_step_control=True, x0='op')),
    'acnl': dict(build=acnl, an=dict(kind='ac', start=1., stop=100e6, points=65, x0='op')),
    'diode_ac': dict(build=diode_ac, an=dict(kind='ac', start=1e3, stop=1e9, points=31, x0='op')),
    'am_src': dict(build=am_src, an=dict(kind='tran', tstart=0., tstop=6e-3, tstep=1e-5,
                                         method='TRAP', use_step_control=True, x0='op')),
    'sffm_src': dict(build=sffm_src, an=dict(kind='tran', tstart=0., tstop=.5e-3, tstep=.5e-6,
                                             method='TRAP', use_step_control=True, x0='op')),
    'exp_src': dict(build=exp_src, an=dict(kind='tran', tstart=0., tstop=6e-4, tstep=1e-6,
                                           method='GEAR2', use_step_control=True, x0='op')),
    'ladder_op': dict(build=diode_ladder, an=dict(kind='op')),
    'ladder_dc': dict(build=diode_ladder, an=dict(kind='dc', start=0., stop=6., points=7, source='V1')),
    'ladder_tran': dict(build=diode_ladder, an=dict(kind='tran', tstart=0., tstop=1e-5, tstep=1e-7,
                                                    method='TRAP', use_step_control=True, x0='op')),
    'ladder_ac': dict(build=diode_ladder, an=dict(kind='ac', start=1e4, stop=1e8, points=9, x0='op')),
    'mesh_op': dict(build=diode_mesh, an=dict(kind='op')),
    'rc_ie': dict(build=rc_gear3, an=dict(kind='tran', tstart=0., tstop=12e-6, tstep=2e-7,
                                          method='IMPLICIT_EULER', use_step_control=True, x0='op')),
}
