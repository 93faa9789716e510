"""Host-side circuit data model: a mirror of the reference's `Circuit` and
element/model classes for the batched Newton/MNA path.

Same public names, argument order and semantics as ahkab/circuit.py:136-1170
(`Circuit(list)`, node numbering, `add_*`, `find_vde_index`, ...),
ahkab/components/*.py, ahkab/diode.py:47-339, ahkab/ekv.py:105-171,367-454,
ahkab/mosq.py:78-137,431-525 and ahkab/switch.py:44-119,272-329.

The classes here only *hold parameters*: no device equation is evaluated on
the host.  Every number is flattened by `compile.py` into the parameter table
that crosses the C-ABI (include/ahkab_b200.h) and is evaluated by the CUDA
engine.  Model constructors keep the reference's derivations (NPMOS sign flip
of VTO, COX from TOX, temperature pre-scaling, ...) written with numpy so that
the same expression maps a scalar *or* a per-instance `(B,)` array.
"""
import math

import numpy as np

from . import constants, options

# element type codes — must match include/ahkab_b200.h
AHK_R, AHK_C, AHK_L, AHK_K, AHK_V, AHK_I, AHK_E, AHK_G, AHK_H, AHK_F, \
    AHK_D, AHK_EKV, AHK_MOSQ, AHK_SW = range(1, 15)

_string_types = (str,)


# --------------------------------------------------------------------------
# linear components (ahkab/components/*.py, ahkab/components/sources/*.py)
# --------------------------------------------------------------------------
class Component(object):
    is_nonlinear = False
    is_symbolic = True
    _ahk_type = 0

    def __repr__(self):
        return "<%s %s>" % (self.__class__.__name__, self.part_id)


class Resistor(Component):
    """ahkab/components/Resistor.py:41-66 (stamps use g = 1/value)."""
    _ahk_type = AHK_R

    def __init__(self, part_id, n1, n2, value):
        self.part_id, self.n1, self.n2, self.value = part_id, n1, n2, value

    @property
    def g(self):
        return 1. / self.value


class Capacitor(Component):
    """ahkab/components/Capacitor.py:46-53."""
    _ahk_type = AHK_C

    def __init__(self, part_id, n1, n2, value, ic=None):
        self.part_id, self.n1, self.n2, self.value, self.ic = \
            part_id, n1, n2, value, ic


class Inductor(Component):
    """ahkab/components/Inductor.py:44-52."""
    _ahk_type = AHK_L

    def __init__(self, part_id, n1, n2, value, ic=None):
        self.part_id, self.n1, self.n2, self.value, self.ic = \
            part_id, n1, n2, value, ic
        self.coupling_devices = []


class InductorCoupling(Component):
    """ahkab/components/InductorCoupling.py; M = K*sqrt(L1*L2)
    (ahkab/circuit.py:606-610) is re-derived per instance on the device."""
    _ahk_type = AHK_K

    def __init__(self, part_id, L1, L2, K, M):
        self.part_id, self.L1, self.L2, self.K, self.M = part_id, L1, L2, K, M

    def get_other_inductor(self, Lselected):
        if Lselected.upper() == self.L1.upper():
            return self.L2
        if Lselected.upper() == self.L2.upper():
            return self.L1
        raise ValueError("Mutual inductors %s and %s coupled by %s" %
                         (self.L1, self.L2, self.part_id))


class _IndepSource(Component):
    def __init__(self, part_id, n1, n2, dc_value, ac_value=0):
        self.part_id, self.n1, self.n2 = part_id, n1, n2
        self.dc_value = dc_value
        # ahkab/components/sources/VSource.py:53-54: falsy ac_value -> None
        self.abs_ac = np.abs(ac_value) if ac_value else None
        self.arg_ac = np.angle(ac_value) if ac_value else None
        self.is_timedependent = False
        self._time_function = None


class VSource(_IndepSource):
    """ahkab/components/sources/VSource.py:48-97."""
    _ahk_type = AHK_V

    def __init__(self, part_id, n1, n2, dc_value, ac_value=0):
        _IndepSource.__init__(self, part_id, n1, n2, dc_value, ac_value)
        if dc_value is not None:
            self.dc_guess = [self.dc_value]


class ISource(_IndepSource):
    """ahkab/components/sources/ISource.py:47-98."""
    _ahk_type = AHK_I


class EVSource(Component):
    """VCVS, ahkab/components/sources/EVSource.py:56-64."""
    _ahk_type = AHK_E

    def __init__(self, part_id, n1, n2, sn1, sn2, value):
        self.part_id, self.n1, self.n2, self.sn1, self.sn2 = \
            part_id, n1, n2, sn1, sn2
        self.alpha = value


class GISource(Component):
    """VCCS, ahkab/components/sources/GISource.py:64-72."""
    _ahk_type = AHK_G

    def __init__(self, part_id, n1, n2, sn1, sn2, value):
        self.part_id, self.n1, self.n2, self.sn1, self.sn2 = \
            part_id, n1, n2, sn1, sn2
        self.alpha = value


class HVSource(Component):
    """CCVS, ahkab/components/sources/HVSource.py:62-69."""
    _ahk_type = AHK_H

    def __init__(self, part_id, n1, n2, source_id, value):
        self.part_id, self.n1, self.n2 = part_id, n1, n2
        self.source_id, self.alpha = source_id, value


class FISource(Component):
    """CCCS, ahkab/components/sources/FISource.py:63-70."""
    _ahk_type = AHK_F

    def __init__(self, part_id, n1, n2, source_id, value):
        self.part_id, self.n1, self.n2 = part_id, n1, n2
        self.source_id, self.alpha = source_id, value


# --------------------------------------------------------------------------
# models: constructor keyword holders + vectorisable derivations
# --------------------------------------------------------------------------
def _f(x):
    """float() for scalars, float64 array for per-instance arrays."""
    if isinstance(x, np.ndarray):
        return x.astype(np.float64)
    return float(x)


class _Model(object):
    """Keeps the constructor keywords so that `runtime(overrides)` can re-run
    the reference's constructor arithmetic with per-instance arrays
    substituted for any keyword (what building a fresh model object per
    instance does in the reference loop)."""
    RUNTIME = ()

    def _store(self, kw):
        self._kw = dict(kw)
        for k, v in self.runtime({}).items():
            setattr(self, k, v)

    def runtime(self, overrides):
        kw = dict(self._kw)
        for k, v in overrides.items():
            if k not in kw:
                raise KeyError("model %s has no parameter %s" % (self.name, k))
            kw[k] = v
        return self._derive(**kw)


class diode_model(_Model):
    """ahkab/diode.py:276-339.  Runtime block (ABI order):
    IS N NBV ISR NR RS BV IKF VT."""
    RUNTIME = ('IS', 'N', 'NBV', 'ISR', 'NR', 'RS', 'BV', 'IKF', 'VT')
    DEFAULTS = dict(IS=1e-14, N=1.0, NBV=1.0, ISR=0.0, NR=2.0, RS=0.0,
                    CJ0=0.0, M=.5, VJ=.7, FC=.5, CP=.0, TT=.0,
                    BV=float('inf'), IBV=1e-3, IKF=float('inf'), KF=.0, AF=1.,
                    FFE=1., TEMP=26.85, XTI=3.0, EG=1.11, TBV=0.0, TRS=0.0,
                    TTT1=0.0, TTT2=0.0, TM1=0.0, TM2=0.0)

    def __init__(self, name, **kw):
        self.name = name
        full = {}
        for k, d in self.DEFAULTS.items():
            v = kw.pop(k, None)
            full[k] = d if v is None else v
        kw.pop('material', None)
        if kw:
            raise TypeError("unknown diode model parameters: %s" % list(kw))
        self._store(full)

    @staticmethod
    def _derive(**kw):
        out = {k: _f(kw[k]) for k in diode_model.RUNTIME if k != 'VT'}
        # T_DEFAULT = Celsius2Kelvin(26.85) (diode.py:272); VT = Vth(T)
        out['VT'] = constants.Vth(26.85 + 273.15)
        return out


class ekv_mos_model(_Model):
    """ahkab/ekv.py:367-454.  Runtime block (ABI order):
    VTO KP GAMMA PHI COX LAMBDA XJ UCRIT; NPMOS travels as an element flag."""
    RUNTIME = ('VTO', 'KP', 'GAMMA', 'PHI', 'COX', 'LAMBDA', 'XJ', 'UCRIT')

    def __init__(self, name=None, TYPE='n', TNOM=None, COX=None, GAMMA=None,
                 NSUB=None, PHI=None, VTO=None, KP=None, XJ=None, LAMBDA=None,
                 TOX=None, VFB=None, U0=None, TCV=None, BEX=None):
        self.name = "model_ekv0" if name is None else name
        self.NPMOS = 1 if TYPE == 'n' else -1
        self._store(dict(TNOM=TNOM, COX=COX, GAMMA=GAMMA, NSUB=NSUB, PHI=PHI,
                         VTO=VTO, KP=KP, XJ=XJ, LAMBDA=LAMBDA, TOX=TOX,
                         VFB=VFB, U0=U0, TCV=TCV, BEX=BEX))
        if not np.all(np.asarray(self.GAMMA) > 0):
            raise Exception("GAMMA " + str(self.GAMMA) + " out of range")
        if not np.all(np.asarray(self.PHI) > 0.1):
            raise Exception("PHI " + str(self.PHI) + " out of range")

    def _derive(self, TNOM, COX, GAMMA, NSUB, PHI, VTO, KP, XJ, LAMBDA, TOX,
                VFB, U0, TCV, BEX):
        NP = self.NPMOS
        TNOM = _f(TNOM) if TNOM is not None else constants.Tref
        if COX is not None:
            COX = _f(COX)
        elif TOX is not None:
            COX = constants.si.eox / TOX
        else:
            COX = .7e-3
        if GAMMA is not None:
            GAMMA = _f(GAMMA)
        elif NSUB is not None:
            GAMMA = np.sqrt(2.0 * constants.e * constants.si.esi * NSUB
                            * 10 ** 6 / COX)
        else:
            GAMMA = 1.0
        if PHI is not None:
            PHI = _f(PHI)
        elif NSUB is not None:
            PHI = 2. * constants.Vth(TNOM) * \
                np.log(NSUB * 10.0 ** 6.0 / constants.si.ni(TNOM))
        else:
            PHI = .7
        if VTO is not None:
            VTO = NP * _f(VTO)
        elif VFB is not None:
            VTO = VFB + PHI + GAMMA * PHI
        else:
            VTO = NP * .5
        if KP is not None:
            KP = _f(KP)
        elif U0 is not None:
            KP = (U0 * 10.0 ** -4) * COX
        else:
            KP = 50e-6
        LAMBDA = LAMBDA if LAMBDA is not None else .5
        XJ = XJ if XJ is not None else .1e-6
        TCV = NP * _f(TCV) if TCV is not None else NP * 1e-3
        BEX = _f(BEX) if BEX is not None else -1.5
        # set_device_temperature(constants.T), ekv.py:444-454
        T = constants.T
        VTO = VTO - TCV * (T - TNOM)
        KP = KP * (T / TNOM) ** BEX
        PHI = PHI * T / TNOM + 3.0 * constants.Vth(TNOM) * math.log(T / TNOM) \
            - constants.si.Eg(TNOM) * T / TNOM + constants.si.Eg(T)
        return dict(VTO=VTO, KP=KP, GAMMA=GAMMA, PHI=PHI, COX=COX,
                    LAMBDA=_f(LAMBDA), XJ=_f(XJ), UCRIT=2e6)


class mosq_mos_model(_Model):
    """ahkab/mosq.py:431-525.  Runtime block (ABI order):
    VTO KP GAMMA PHI LAMBDA AVT AKP."""
    RUNTIME = ('VTO', 'KP', 'GAMMA', 'PHI', 'LAMBDA', 'AVT', 'AKP')

    def __init__(self, name=None, TYPE='n', TNOM=None, COX=None, GAMMA=None,
                 NSUB=None, PHI=None, VTO=None, KP=None, LAMBDA=None,
                 AKP=None, AVT=None, TOX=None, VFB=None, U0=None, TCV=None,
                 BEX=None):
        self.name = "model_mosq0" if name is None else name
        if TYPE.lower() == 'n':
            self.NPMOS = 1
        elif TYPE.lower() == 'p':
            self.NPMOS = -1
        else:
            raise ValueError("Unknown MOS type %s" % TYPE)
        self._store(dict(TNOM=TNOM, COX=COX, GAMMA=GAMMA, NSUB=NSUB, PHI=PHI,
                         VTO=VTO, KP=KP, LAMBDA=LAMBDA, AKP=AKP, AVT=AVT,
                         TOX=TOX, VFB=VFB, U0=U0, TCV=TCV, BEX=BEX))
        if not np.all(np.asarray(self.GAMMA) > 0):
            raise Exception("GAMMA " + str(self.GAMMA) + " out of range")
        if not np.all(np.asarray(self.PHI) > 0.1):
            raise Exception("PHI " + str(self.PHI) + " out of range")

    def _derive(self, TNOM, COX, GAMMA, NSUB, PHI, VTO, KP, LAMBDA, AKP, AVT,
                TOX, VFB, U0, TCV, BEX):
        NP = self.NPMOS
        TNOM = _f(TNOM) if TNOM is not None else constants.Tref
        if COX is not None:
            COX = _f(COX)
        elif TOX is not None:
            COX = constants.si.eox / TOX
        else:
            COX = .7e-3
        if GAMMA is not None:
            GAMMA = _f(GAMMA)
        elif NSUB is not None:
            GAMMA = np.sqrt(2 * constants.e * constants.si.esi * NSUB
                            * 10 ** 6 / COX)
        else:
            GAMMA = 1.0
        if PHI is not None:
            PHI = _f(PHI)
        elif NSUB is not None:
            PHI = 2 * constants.Vth(TNOM) * np.log(
                NSUB * 10 ** 6 / constants.si.ni(TNOM))
        else:
            PHI = .7
        if VTO is not None:
            VTO = NP * _f(VTO)
        elif VFB is not None:
            VTO = VFB + PHI + GAMMA * PHI
        else:
            VTO = .5            # mosq.py:482: *not* sign-flipped
        if KP is not None:
            KP = _f(KP)
        elif U0 is not None:
            KP = (U0 * 10 ** -4) * COX
        else:
            KP = 50e-6
        LAMBDA = LAMBDA if LAMBDA is not None else .5
        TCV = NP * _f(TCV) if TCV is not None else NP * 1e-3
        BEX = _f(BEX) if BEX is not None else -1.5
        AVT = AVT if AVT is not None else 7.1e-3 * 1e-6
        AKP = AKP if AKP is not None else 1.8e-2 * 1e-6
        T = constants.T
        VTO = VTO - TCV * (T - TNOM)
        KP = KP * (T / TNOM) ** BEX
        PHI = (PHI * T / TNOM + 3 * constants.Vth(TNOM) * math.log(T / TNOM)
               - constants.si.Eg(TNOM) * T / TNOM + constants.si.Eg(T))
        return dict(VTO=VTO, KP=KP, GAMMA=GAMMA, PHI=PHI, LAMBDA=_f(LAMBDA),
                    AVT=_f(AVT), AKP=_f(AKP))


class vswitch_model(_Model):
    """ahkab/switch.py:272-329.  Runtime block (ABI order): VT VH RON ROFF."""
    RUNTIME = ('VT', 'VH', 'RON', 'ROFF')

    def __init__(self, name, VT=None, VH=None, VON=None, VOFF=None, RON=None,
                 ROFF=None):
        self.name = name
        self._store(dict(VT=VT, VH=VH, VON=VON, VOFF=VOFF, RON=RON, ROFF=ROFF))

    @staticmethod
    def _derive(VT, VH, VON, VOFF, RON, ROFF):
        if VON is not None or VOFF is not None:
            VON, VOFF = _f(VON), _f(VOFF)
            VT = (VON - VOFF) / 2.0 + VOFF
            VH = VT * 1e-3
        VT = _f(VT) if VT is not None else 0.0
        VH = _f(VH) if VH is not None else 0.0
        RON = _f(RON) if RON is not None else 1.
        ROFF = _f(ROFF) if ROFF is not None else 1. / options.gmin
        return dict(VT=VT, VH=VH, RON=RON, ROFF=ROFF)

    def get_dc_guess(self, is_on):
        return [self.VT * (.9 + is_on * .2)] * 2


# --------------------------------------------------------------------------
# non-linear devices (parameter holders only)
# --------------------------------------------------------------------------
class _DevParams(object):
    pass


class diode(Component):
    """ahkab/diode.py:47-104."""
    is_nonlinear = True
    _ahk_type = AHK_D

    def __init__(self, part_id, n1, n2, model, AREA=None, T=None, ic=None,
                 off=False):
        self.part_id, self.n1, self.n2, self.model = part_id, n1, n2, model
        self.device = _DevParams()
        self.device.AREA = AREA if AREA is not None else 1.0
        self.device.T = T if T is not None else constants.T
        if self.device.T != constants.T:
            raise NotImplementedError(
                "diode %s: per-device temperature != %g K is not supported by "
                "the batched engine" % (part_id, constants.T))
        self.ports = ((self.n1, self.n2),)
        self.dc_guess = [0.425]
        if ic is not None:
            self.dc_guess = ic
        self.ic = ic
        self.off = off
        if self.off and self.ic is None:
            self.ic = 0

    def get_output_ports(self):
        return self.ports

    def get_drive_ports(self, op):
        return self.ports


class ekv_device(Component):
    """ahkab/ekv.py:105-171.  Drive ports (d,b),(g,b),(s,b); output (d,s)."""
    is_nonlinear = True
    _ahk_type = AHK_EKV

    def __init__(self, part_id, nd, ng, ns, nb, W, L, model, M=1, N=1):
        self.part_id = part_id
        self.ng, self.nb, self.n1, self.n2 = ng, nb, nd, ns
        self.ports = ((self.n1, self.nb), (self.ng, self.nb),
                      (self.n2, self.nb))
        self.device = _DevParams()
        self.device.L, self.device.W = float(L), float(W)
        self.device.M, self.device.N = int(M), int(N)
        self.ekv_model = self.model = model
        for nm in ('L', 'W', 'N', 'M'):
            if not getattr(self.device, nm) > 0:
                raise Exception(nm + " out of boundaries.")

    @property
    def dc_guess(self):
        m = self.model
        return [m.VTO * (0.1) * m.NPMOS, m.VTO * (1.1) * m.NPMOS, 0]

    def get_drive_ports(self, op):
        return self.ports

    def get_output_ports(self):
        return ((self.n1, self.n2),)


class mosq_device(Component):
    """ahkab/mosq.py:78-137.  Drive ports (d,s),(g,s),(b,s); output (d,s)."""
    is_nonlinear = True
    _ahk_type = AHK_MOSQ

    def __init__(self, part_id, nd, ng, ns, nb, W, L, model, M=1, N=1):
        self.part_id = part_id
        self.ng, self.nb, self.n1, self.n2 = ng, nb, nd, ns
        self.ports = ((self.n1, self.n2), (self.ng, self.n2),
                      (self.nb, self.n2))
        self.device = _DevParams()
        self.device.L, self.device.W = float(L), float(W)
        self.device.M, self.device.N = int(M), int(N)
        self.device.mckey = None
        self.mosq_model = self.model = model
        for nm in ('L', 'W', 'N', 'M'):
            if not getattr(self.device, nm) > 0:
                raise ValueError(nm + " out of boundaries.")

    @property
    def dc_guess(self):
        m = self.model
        return [m.VTO * 0.4 * m.NPMOS, m.VTO * 1.1 * m.NPMOS, 0]

    def setup_mc(self, status, mckey):
        """Pelgrom mismatch hook, ahkab/mosq.py:411-419."""
        self.device.mckey = mckey if status else None

    def get_drive_ports(self, op):
        return self.ports

    def get_output_ports(self):
        return ((self.n1, self.n2),)


class switch_device(Component):
    """ahkab/switch.py:44-119.  Ports (n1,n2),(sn1,sn2)."""
    is_nonlinear = True
    _ahk_type = AHK_SW

    def __init__(self, n1, n2, sn1, sn2, model, ic=None, part_id='S'):
        self.device = _DevParams()
        self.device.is_on = ic if ic is not None else False
        self.sn1, self.sn2, self.n1, self.n2 = sn1, sn2, n1, n2
        self.ports = ((self.n1, self.n2), (self.sn1, self.sn2))
        self.model = model
        self.part_id = part_id

    @property
    def dc_guess(self):
        return self.model.get_dc_guess(self.device.is_on)

    def get_drive_ports(self, op):
        return self.ports

    def get_output_ports(self):
        return ((self.n1, self.n2),)


def is_elem_voltage_defined(elem):
    """ahkab/circuit.py:1150-1169."""
    return isinstance(elem, (VSource, EVSource, HVSource, Inductor))


# --------------------------------------------------------------------------
# Circuit
# --------------------------------------------------------------------------
class Circuit(list):
    """A list of elements plus the node maps — ahkab/circuit.py:136-1147."""

    def __init__(self, title, filename=None):
        list.__init__(self)
        self.title = title
        self.filename = filename
        self.nodes_dict = {}
        self.internal_nodes = 0
        self.models = {}
        self.gnd = '0'

    # -- nodes (circuit.py:203-290) --------------------------------------
    def create_node(self, name):
        if type(name) not in _string_types:
            raise TypeError("The node %s should have been of text type" % name)
        got_ref = 0 in self.nodes_dict
        if name not in self.nodes_dict:
            if name == '0':
                int_node = 0
            else:
                int_node = int(len(self.nodes_dict) / 2) + 1 * (not got_ref)
            self.nodes_dict.update({int_node: name})
            self.nodes_dict.update({name: int_node})
        else:
            raise ValueError('Impossible to create new node %s: node exists!'
                             % name)
        return name

    def add_node(self, ext_name):
        if type(ext_name) not in _string_types:
            raise TypeError("The node %s should have been of text type" %
                            ext_name)
        if ext_name not in self.nodes_dict:
            if ext_name == '0':
                int_node = 0
            else:
                got_ref = 0 in self.nodes_dict
                int_node = int(len(self.nodes_dict) / 2) + 1 * (not got_ref)
            self.nodes_dict.update({int_node: ext_name})
            self.nodes_dict.update({ext_name: int_node})
        else:
            int_node = self.nodes_dict[ext_name]
        return int_node

    def new_internal_node(self):
        ext_node = "INT" + str(self.internal_nodes)
        self.internal_nodes = self.internal_nodes + 1
        self.create_node(ext_node)
        return ext_node

    def get_nodes_number(self):
        return int(len(self.nodes_dict) / 2)

    def is_nonlinear(self):
        for elem in self:
            if elem.is_nonlinear:
                return True
        return False

    def get_locked_nodes(self):
        """circuit.py:316-336: every drive port of every non-linear element."""
        locked_nodes = []
        for elem in self:
            if not elem.is_nonlinear:
                continue
            oports = elem.get_output_ports()
            for index in range(len(oports)):
                for port in elem.get_drive_ports(index):
                    locked_nodes.append(port)
        return locked_nodes

    def ext_node_to_int(self, ext_node):
        return self.nodes_dict[ext_node]

    def int_node_to_ext(self, int_node):
        return self.nodes_dict[int_node]

    def has_duplicate_elem(self):
        all_ids = tuple(map(lambda e: e.part_id, self))
        return len(set(all_ids)) != len(all_ids)

    def get_ground_node(self):
        return '0'

    def get_elem_by_name(self, part_id):
        for e in self:
            if e.part_id.lower() == part_id.lower():
                return e
        raise ValueError('Element %s not found' % part_id)

    # -- models (circuit.py:379-409) ---------------------------------------
    def add_model(self, model_type, model_label, model_parameters):
        if 'name' not in model_parameters:
            model_parameters.update({'name': model_label})
        if model_type == "ekv":
            model_iter = ekv_mos_model(**model_parameters)
        elif model_type == "mosq":
            model_iter = mosq_mos_model(**model_parameters)
        elif model_type == "diode":
            model_iter = diode_model(**model_parameters)
        elif model_type == "sw":
            model_iter = vswitch_model(**model_parameters)
        else:
            raise ValueError("Unknown model %s" % (model_type,))
        self.models.update({model_label: model_iter})

    def remove_model(self, model_label):
        if self.models is not None and model_label in self.models:
            del self.models[model_label]

    # -- elements (circuit.py:411-971) ----------------------------------------
    def add_resistor(self, part_id, n1, n2, value):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        self.append(Resistor(part_id=part_id, n1=n1, n2=n2, value=value))

    def add_capacitor(self, part_id, n1, n2, value, ic=None):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        self.append(Capacitor(part_id=part_id, n1=n1, n2=n2, value=value,
                              ic=ic))

    def add_inductor(self, part_id, n1, n2, value, ic=None):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        self.append(Inductor(part_id=part_id, n1=n1, n2=n2, value=value,
                             ic=ic))

    def add_inductor_coupling(self, part_id, L1, L2, value):
        L1elem, L2elem = None, None
        for e in self:
            if isinstance(e, Inductor) and (L1 == e.part_id):
                L1elem = e
            elif isinstance(e, Inductor) and (L2 == e.part_id):
                L2elem = e
        if L1elem is None or L2elem is None:
            raise ValueError("One or more coupled inductors for %s were not "
                             "found: %s (found: %s), %s (found: %s)." %
                             (part_id, L1, L1elem is not None, L2,
                              L2elem is not None))
        M = math.sqrt(L1elem.value * L2elem.value) * value
        elem = InductorCoupling(part_id=part_id, L1=L1, L2=L2, K=value, M=M)
        L1elem.coupling_devices.append(elem)
        L2elem.coupling_devices.append(elem)
        self.append(elem)

    def add_vsource(self, part_id, n1, n2, dc_value, ac_value=0,
                    function=None):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        elem = VSource(part_id=part_id, n1=n1, n2=n2, dc_value=dc_value,
                       ac_value=ac_value)
        if function is not None:
            elem.is_timedependent = True
            elem._time_function = function
        self.append(elem)

    def add_isource(self, part_id, n1, n2, dc_value, ac_value=0,
                    function=None):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        elem = ISource(part_id=part_id, n1=n1, n2=n2, dc_value=dc_value,
                       ac_value=ac_value)
        if function is not None:
            elem.is_timedependent = True
            elem._time_function = function
        self.append(elem)

    def add_diode(self, part_id, n1, n2, model_label, models=None, Area=None,
                  T=None, ic=None, off=False):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        if models is None:
            models = self.models
        if model_label not in models:
            raise ValueError("Unknown diode model id: " + model_label)
        self.append(diode(part_id=part_id, n1=n1, n2=n2,
                          model=models[model_label], AREA=Area, T=T, ic=ic,
                          off=off))

    def add_mos(self, part_id, nd, ng, ns, nb, w, l, model_label, models=None,
                m=1, n=1):
        nd, ng = self.add_node(nd), self.add_node(ng)
        ns, nb = self.add_node(ns), self.add_node(nb)
        if models is None:
            models = self.models
        if model_label not in models:
            raise ValueError("Unknown model id: " + model_label)
        if isinstance(models[model_label], ekv_mos_model):
            elem = ekv_device(part_id, nd, ng, ns, nb, w, l,
                              models[model_label], m, n)
        elif isinstance(models[model_label], mosq_mos_model):
            elem = mosq_device(part_id, nd, ng, ns, nb, w, l,
                               models[model_label], m, n)
        else:
            raise Exception("Unknown model type for " + model_label)
        self.append(elem)

    def add_cccs(self, part_id, n1, n2, source_id, value):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        self.append(FISource(part_id=part_id, n1=n1, n2=n2,
                             source_id=source_id, value=value))

    def add_ccvs(self, part_id, n1, n2, source_id, value):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        self.append(HVSource(part_id=part_id, n1=n1, n2=n2,
                             source_id=source_id, value=value))

    def add_vcvs(self, part_id, n1, n2, sn1, sn2, value):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        sn1, sn2 = self.add_node(sn1), self.add_node(sn2)
        self.append(EVSource(part_id=part_id, n1=n1, n2=n2, sn1=sn1, sn2=sn2,
                             value=value))

    def add_vccs(self, part_id, n1, n2, sn1, sn2, value):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        sn1, sn2 = self.add_node(sn1), self.add_node(sn2)
        self.append(GISource(part_id=part_id, n1=n1, n2=n2, sn1=sn1, sn2=sn2,
                             value=value))

    def add_switch(self, part_id, n1, n2, sn1, sn2, ic, model_label,
                   models=None):
        n1, n2 = self.add_node(n1), self.add_node(n2)
        sn1, sn2 = self.add_node(sn1), self.add_node(sn2)
        if models is None:
            models = self.models
        if model_label not in models:
            raise ValueError("Unknown switch model id: " + model_label)
        self.append(switch_device(n1=n1, n2=n2, sn1=sn1, sn2=sn2,
                                  model=models[model_label], ic=ic,
                                  part_id=part_id))

    def add_user_defined(self, module, label, param_dict):
        raise NotImplementedError(
            "user-defined Python elements cannot run on the device "
            "(SURVEY.md §2 row 29, out of scope)")

    # -- voltage-defined element indexing (circuit.py:1078-1147) -------------
    def find_vde_index(self, elem_or_id, verbose=3):
        if hasattr(elem_or_id, 'part_id'):
            part_id = elem_or_id.part_id
        else:
            part_id = elem_or_id
        part_id = part_id.lower()
        vde_index = 0
        found = False
        for e in self:
            if is_elem_voltage_defined(e):
                if e.part_id.lower() == part_id:
                    found = True
                    break
                vde_index += 1
        if not found:
            raise ValueError("find_vde_index(): element %s was not found."
                             % (part_id,))
        return vde_index

    def find_vde(self, index):
        index = index - self.get_nodes_number() + 1
        ni = 0
        for e in self:
            if is_elem_voltage_defined(e):
                if index == ni:
                    return e
                ni = ni + 1
        raise IndexError('no voltage defined element with index %d' % index)
