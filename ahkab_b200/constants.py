"""Physical constants, same values and expressions as the reference
(ahkab/constants.py:30-53, 91-112) so that derived quantities round identically."""
import math

e = 1.60217646e-19
T = 300
Tref = 300
k = 1.3806503e-23


def Vth(T=Tref):
    """Thermal voltage kT/q (ahkab/constants.py:36-52)."""
    return k * T / e


class _Silicon(object):
    """ahkab/constants.py:55-112."""
    def __init__(self):
        self.esi = 104.5 * 10 ** -12
        self.eox = 34.5 * 10 ** -12
        self.Eg_0 = 1.16
        self.alpha = 0.000702
        self.beta = 1108
        self.n_i_0 = 1.45 * 10 ** 16

    def Eg(self, T=Tref):
        return (self.Eg_0 - self.alpha * T ** 2 / (self.beta + T))

    def ni(self, T=Tref):
        return self.n_i_0 * (T / Tref) ** (3 / 2) * math.exp(
            self.Eg(Tref) / (2 * Vth(Tref)) - self.Eg(T) / (2 * Vth(T)))


si = _Silicon()
