"""Time functions of independent sources — parameter holders mirroring
ahkab/time_functions.py:402-745 (pulse, sin, exp, sffm, am).

The classes keep the reference's constructor signatures.  `__call__` is
provided for host-side inspection (plots, tests); during a transient the
waveform is evaluated on the device from the 8-slot parameter block
`encode()` produces (include/ahkab_b200.h: AHK_TF_*), so the batch axis can
also vary waveform parameters (`pwl` carries its table: 4 + 2*npts slots)."""
import math

TF_NONE, TF_PULSE, TF_SIN, TF_EXP, TF_SFFM, TF_AM, TF_PWL = range(7)


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


class pwl(object):
    """ahkab/time_functions.py:730-828: piece-wise linear through (x_i, y_i) — a k = 1
    InterpolatedUnivariateSpline, i.e. linear extrapolation outside the table — with an
    optional delay and repetition."""
    def __init__(self, x, y, repeat=False, repeat_time=0, td=0):
        self.x = [float(v) for v in x]
        self.y = [float(v) for v in y]
        if len(self.x) != len(self.y) or len(self.x) < 2:
            raise ValueError("pwl needs two sequences of at least 2 points")
        self.repeat = repeat
        self.repeat_time = repeat_time
        if self.repeat_time == max(self.x):
            self.repeat_time = 0
        self.td = td
        self._type = "V"

    def _normalize_time(self, time):
        if time is None:
            time = 0
        if time <= self.td:
            time = 0
        else:
            time = time - self.td
            if self.repeat and time > max(self.x):
                time = (time - max(self.x)) % (max(self.x) - self.repeat_time) + \
                    self.repeat_time
        return time

    def __call__(self, time):
        t = self._normalize_time(time)
        i = 0
        while i < len(self.x) - 2 and t > self.x[i + 1]:
            i += 1
        x0, y0, x1, y1 = self.x[i], self.y[i], self.x[i + 1], self.y[i + 1]
        return y0 + (y1 - y0) / (x1 - x0) * (t - x0)

    def _encode(self):
        xy = []
        for a, b in zip(self.x, self.y):
            xy += [a, b]
        return TF_PWL, [float(len(self.x)), self.td, 1.0 if self.repeat else 0.0,
                        self.repeat_time] + xy


def encode(fn):
    """(kind, params) for the ABI parameter block: up to 8 values, or 4 + 2*npts for pwl."""
    if fn is None:
        return TF_NONE, []
    if hasattr(fn, '_encode'):
        return fn._encode()
    raise NotImplementedError(
        "time function %r cannot be evaluated on the device: only pulse, sin, "
        "exp, sffm, am and pwl are carried across the ABI" % (fn,))
