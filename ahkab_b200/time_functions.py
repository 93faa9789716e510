"""Time functions of independent sources — parameter holders mirroring
ahkab/time_functions.py:402-745 (pulse, sin, exp, sffm, am).

The classes keep the reference's constructor signatures.  `__call__` is
provided for host-side inspection (plots, tests); during a transient the
waveform is evaluated on the device from the 8-slot parameter block
`encode()` produces (include/ahkab_b200.h: AHK_TF_*), so the batch axis can
also vary waveform parameters.  `pwl` (variable-length spline table) is not
carried across the ABI in this round."""
import math

TF_NONE, TF_PULSE, TF_SIN, TF_EXP, TF_SFFM, TF_AM = range(6)


class pulse(object):
    """ahkab/time_functions.py:402-470."""
    def __init__(self, v1, v2, td, tr, pw, tf, per):
        self.v1, self.v2 = v1, v2
        self.td = max(td, 0.0)
        self.per, self.tr, self.tf, self.pw = per, tr, tf, pw
        self._type = "V"

    def __call__(self, time):
        if time is None:
            time = 0
        time = time - self.per * int(time / self.per)
        if time < self.td:
            return self.v1
        elif time < self.td + self.tr:
            return self.v1 + ((self.v2 - self.v1) / (self.tr)) * \
                (time - self.td)
        elif time < self.td + self.tr + self.pw:
            return self.v2
        elif time < self.td + self.tr + self.pw + self.tf:
            return self.v2 + ((self.v1 - self.v2) / (self.tf)) * \
                (time - (self.td + self.tr + self.pw))
        else:
            return self.v1

    def _encode(self):
        return TF_PULSE, [self.v1, self.v2, self.td, self.tr, self.pw,
                          self.tf, self.per]


class sin(object):
    """ahkab/time_functions.py:473-545."""
    def __init__(self, vo, va, freq, td=0., theta=0., phi=0.):
        self.vo, self.va, self.freq = vo, va, freq
        self.td, self.theta, self.phi = td, theta, phi
        self._type = "V"

    def __call__(self, time):
        if time is None:
            time = 0
        if time < self.td:
            return self.vo + self.va * math.sin(math.pi * self.phi / 180.)
        return self.vo + self.va * math.exp((self.td - time) * self.theta) \
            * math.sin(2 * math.pi * self.freq * (time - self.td) +
                       math.pi * self.phi / 180.)

    def _encode(self):
        return TF_SIN, [self.vo, self.va, self.freq, self.td, self.theta,
                        self.phi]


class exp(object):
    """ahkab/time_functions.py:548-618."""
    def __init__(self, v1, v2, td1, tau1, td2, tau2):
        self.v1, self.v2, self.td1 = v1, v2, td1
        self.tau1, self.td2, self.tau2 = tau1, td2, tau2
        self._type = "V"

    def __call__(self, time):
        if time is None:
            time = 0
        if time < self.td1:
            return self.v1
        elif time < self.td2:
            return self.v1 + (self.v2 - self.v1) * \
                (1 - math.exp(-1 * (time - self.td1) / self.tau1))
        return self.v1 + (self.v2 - self.v1) * \
            (1 - math.exp(-1 * (time - self.td1) / self.tau1)) + \
            (self.v1 - self.v2) * \
            (1 - math.exp(-1 * (time - self.td2) / self.tau2))

    def _encode(self):
        return TF_EXP, [self.v1, self.v2, self.td1, self.tau1, self.td2,
                        self.tau2]


class sffm(object):
    """ahkab/time_functions.py:621-680."""
    def __init__(self, vo, va, fc, mdi, fs, td):
        self.vo, self.va, self.fc = vo, va, fc
        self.mdi, self.fs, self.td = mdi, fs, td
        self._type = "V"

    def __call__(self, time):
        if time is None:
            time = 0
        if time <= self.td:
            return self.vo
        return self.vo + self.va * math.sin(
            2 * math.pi * self.fc * (time - self.td) +
            self.mdi * math.sin(2 * math.pi * self.fs * (time - self.td)))

    def _encode(self):
        return TF_SFFM, [self.vo, self.va, self.fc, self.mdi, self.fs,
                         self.td]


class am(object):
    """ahkab/time_functions.py:683-745."""
    def __init__(self, sa, fc, fm, oc, td):
        self.sa, self.fc, self.fm, self.oc, self.td = sa, fc, fm, oc, td
        self._type = "V"

    def __call__(self, time):
        if time is None:
            time = 0
        if time <= self.td:
            return 0.
        return self.sa * (self.oc + math.sin(2 * math.pi * self.fm *
                                             (time - self.td))) * \
            math.sin(2 * math.pi * self.fc * (time - self.td))

    def _encode(self):
        return TF_AM, [self.sa, self.fc, self.fm, self.oc, self.td]


def encode(fn):
    """(kind, params[<=8]) for the ABI parameter block."""
    if fn is None:
        return TF_NONE, []
    if hasattr(fn, '_encode'):
        return fn._encode()
    raise NotImplementedError(
        "time function %r cannot be evaluated on the device: only pulse, sin, "
        "exp, sffm and am are carried across the ABI" % (fn,))
