"""Results objects with a leading batch axis.

Same variable names, ordering and dict-like access as the reference's
`op_solution`, `dc_solution`, `tran_solution`, `ac_solution`
(ahkab/results.py:245-853): `V<node>` for the internal nodes 1..nv-1 in creation
order, then `I(<PART_ID>)` for the voltage-defined elements in circuit order;
OP/TRAN names are upper-cased, DC/AC keep the node's case; lookups are
case-insensitive.  The reference streams every row to a tab-separated text
file and re-parses it on each access (results.py:121-124,178-186); here the
data stays in tensors (device until first host access) — no per-row file I/O.

Shapes:  op[var] -> (B,)   dc[var] -> (B, npts)   tran[var] -> (B, max_rows)
(rows beyond `rows[b]` are NaN)   ac[var] -> (B, npts) complex128."""
import numpy as np

from . import circuit as _c
from . import _abi


def _varnames(circ, upper):
    nv_1 = circ.get_nodes_number() - 1
    names, units = [], {}
    for index in range(nv_1):
        v = "V%s" % (str(circ.nodes_dict[index + 1]),)
        v = v.upper() if upper else v
        names.append(v)
        units[v] = "V"
    for elem in circ:
        if _c.is_elem_voltage_defined(elem):
            v = "I(%s)" % (elem.part_id.upper(),)
            v = v.upper() if upper else v
            names.append(v)
            units[v] = "A"
    return names, units


def _host(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t
    return t.detach().cpu().numpy()


class solution(object):
    """Base: dict-like access by variable name (results.py:127-243)."""
    sol_type = None

    def __init__(self, circ, raw, engine=None):
        self.netlist_title = circ.title
        self.raw = raw          # dict of (device) tensors from the engine
        self._engine = engine   # (for the one-copy host read of the packed result set)
        self._np = {}
        self.variables = []
        self.units = {}

    def _get(self, key):
        if key not in self._np:
            pk = self.raw.get('_packed') if isinstance(self.raw, dict) else None
            if self._engine is not None and pk is not None and not self._np.get('_fetched'):
                # first host access: every tensor of the call in one device->host copy
                # (Engine.to_host_owned; the arrays own their pinned buffer)
                host = self._engine.to_host_owned(self.raw)
                self._np['_fetched'] = True
                for k, v in host.items():
                    rv = self.raw.get(k)
                    if rv is not None and hasattr(rv, 'is_complex') and rv.is_complex():
                        v = v.view(np.complex128).reshape(tuple(rv.shape))
                    self._np.setdefault(k, v)
            if key not in self._np:
                self._np[key] = _host(self.raw[key])
        return self._np[key]

    def _index(self, name):
        up = [v.upper() for v in self.variables]
        try:
            return up.index(name.upper())
        except ValueError:
            raise KeyError(name)

    def __len__(self):
        return len(self.variables)

    def __contains__(self, name):
        return name.upper() in [v.upper() for v in self.variables]

    has_key = __contains__

    def keys(self):
        return list(self.variables)

    def values(self):
        return [self[v] for v in self.variables]

    def items(self):
        return list(zip(self.keys(), self.values()))

    def __iter__(self):
        return iter(self.items())

    def get(self, name, default=None):
        try:
            return self[name]
        except KeyError:
            return default

    @property
    def status(self):
        return self._get('status')

    def instance_array(self, b):
        """(variables, samples) array of instance b — what the reference's
        `solution.asarray()` returns for its single circuit."""
        a = self.asarray()
        return np.asarray(a[b]).reshape(len(self.variables), -1)

    def write_csv(self, filename, b=0):
        """Optional exporter: instance b in the reference's on-disk results format
        (csvlib.write_csv, ahkab/csvlib.py:64-100): '#'-prefixed tab-separated header,
        one row per sample, numpy.savetxt's '%.18e' (lossless for float64).  AC files
        hold |x| and arg(x) columns like the reference (results.py:640-682)."""
        data = self.instance_array(b)
        headers = list(self.variables)
        if np.iscomplexobj(data):
            rows, names = [np.real(data[0])], [headers[0]]
            for i in range(1, len(headers)):
                rows += [np.abs(data[i]), np.angle(data[i])]
                names += ['|%s|' % headers[i], 'arg(%s)' % headers[i]]
            data, headers = np.stack(rows), names
        with open(filename, 'wb') as fp:
            np.savetxt(fp, data.T, delimiter=u"\t", header=u"\t".join(headers), comments='#')

    @property
    def ok(self):
        return (self.status & _abi.ST_OK) != 0


class op_solution(solution):
    """results.py:245-551 with a batch axis: sol['VN1'] -> (B,)."""
    sol_type = "OP"

    def __init__(self, circ, raw, engine=None):
        solution.__init__(self, circ, raw, engine)
        self.variables, self.units = _varnames(circ, upper=True)

    @property
    def iterations(self):
        return self._get('iters')

    @property
    def x(self):
        return self._get('x')

    @property
    def errors(self):
        return self._get('resid')

    def asarray(self):
        """(B, n) — the reference returns (n, 1) for its single instance."""
        return self.x

    def __getitem__(self, name):
        return self.x[:, self._index(name)]

    def get_table_array(self, b=0):
        """results.py:325-336: the RESULTS table of instance b (Variable, Units, Value, Error, %)."""
        import tabulate
        x, err = self.x[b], self.errors[b]
        table = []
        for i, v in enumerate(self.variables):
            rel = err[i] / x[i] * 100.0 if x[i] != 0 else 0.0
            table.append((v, self.units[v], float(x[i]), float(err[i]),
                          '%d' % rel if np.isfinite(rel) else 'nan'))
        return tabulate.tabulate(table, headers=("Variable", "Units", "Value", "Error", "%"))

    def _elements_op(self, circ, b):
        """results.py:338-438 for instance b: the per-element OP tables of the linear elements
        (resistors, independent sources — `get_op_info` of Resistor.py:71-92, VSource.py:126-149,
        ISource.py) and the total power delivered by the sources.  Device reports (diode / MOS
        small-signal tables) are not produced: they re-evaluate the Python device models."""
        x = self.x[b]
        nv_1 = circ.get_nodes_number() - 1
        vde = [e for e in circ if _c.is_elem_voltage_defined(e)]

        def vof(n1, n2):
            return (x[n1 - 1] if n1 else 0.0) - (x[n2 - 1] if n2 else 0.0)
        info, tot_power = {}, 0.0
        for e in circ:
            pid = e.part_id.upper()
            if isinstance(e, _c.Resistor):
                v = vof(e.n1, e.n2)
                info.setdefault('R', (['Part ID', u"R [\u2126]", "V(n1,n2) [V]", "I(n1->n2) [A]", "P [W]"], []))[1] \
                    .append([pid, e.value, float(v), float(v / e.value), float(v * v / e.value)])
            elif isinstance(e, _c.VSource):
                i = float(x[nv_1 + vde.index(e)])
                V = e.dc_value if e.dc_value is not None else 0.0
                info.setdefault('V', (['Part ID', "V(n1,n2) [V]", "I(n1->n2) [A]", "P [W]"], []))[1] \
                    .append([pid, V, i, V * i])
                tot_power -= vof(e.n1, e.n2) * i
            elif isinstance(e, _c.ISource):
                v, I = float(vof(e.n1, e.n2)), (e.dc_value if e.dc_value is not None else 0.0)
                info.setdefault('I', (['Part ID', "V(n1-n2) [V]", "I [A]", "P [W]"], []))[1] \
                    .append([pid, v, I, v * I])
                tot_power -= v * I
            elif isinstance(e, _c.EVSource):
                tot_power -= vof(e.n1, e.n2) * float(x[nv_1 + vde.index(e)])
            elif isinstance(e, _c.GISource):
                tot_power -= vof(e.n1, e.n2) * vof(e.sn1, e.sn2) * e.alpha
        return info, tot_power

    def write_opinfo(self, filename, circ, b=0, options=None):
        """Optional exporter: the reference's human-readable `.opinfo` report for instance b
        (op_solution.write_to_file, ahkab/results.py:441-475): banner, options, iteration count,
        the RESULTS table, the element OP tables and the total power dissipation."""
        import time
        import tabulate
        from . import options as _o
        o = _o.snapshot()
        if options:
            o.update(options)
        info, tot_power = self._elements_op(circ, b)
        with open(filename, 'w', encoding='utf-8') as fp:
            fp.write(time.strftime("%c") + "\n")
            fp.write("ahkab_b200 (batched hot path of ahkab v0.18), instance %d of %d\n\n" % (b, len(self.x)))
            fp.write("Operating Point (OP) analysis\n\n")
            fp.write("Netlist: %s\nTitle: %s\n" % (None, self.netlist_title))
            fp.write("At %.2f K\n" % (300.0,))
            fp.write("Options:\n\tvea = %e\n\tver = %f\n\tiea = %e\n\tier = %f\n\tgmin = %e\n"
                     % (o['vea'], o['ver'], o['iea'], o['ier'], o['gmin']))
            fp.write("\nConvergence reached in %d iterations.\n" % (int(self.iterations[b]),))
            fp.write("\n========\nRESULTS:\n========\n\n")
            fp.write(self.get_table_array(b) + '\n')
            fp.write("\n========================\nELEMENTS OP INFORMATION:\n========================\n\n")
            for k in sorted(info):
                fp.write(tabulate.tabulate(info[k][1], headers=info[k][0]) + '\n\n')
            fp.write('Total power dissipation: %g W\n\n' % tot_power)

    @property
    def gmin_check_failed(self):
        return (self.status & _abi.ST_GMIN_CHECK) != 0


class dc_solution(solution):
    """results.py:714-782: variables[0] is the sweep variable."""
    sol_type = "DC"

    def __init__(self, circ, raw, sweepvar, stype, engine=None):
        solution.__init__(self, circ, raw, engine)
        names, units = _varnames(circ, upper=False)
        self.variables = [sweepvar] + names
        self.units = dict(units)
        self.units[sweepvar] = 'V' if sweepvar[0] == 'V' else 'A'
        self.stype = stype

    def asarray(self):
        """(B, 1 + n, npts): sweep axis first, like the reference's rows."""
        x = self._get('x')
        sw = np.broadcast_to(self.raw['sweep'][None, None, :],
                             (x.shape[0], 1, x.shape[1]))
        return np.concatenate((sw, np.transpose(x, (0, 2, 1))), axis=1)

    def __getitem__(self, name):
        i = self._index(name)
        if i == 0:
            return self.raw['sweep']
        return self._get('x')[:, :, i - 1]

    def get_x(self):
        return self.raw['sweep']

    def get_xlabel(self):
        return self.variables[0]


class tran_solution(solution):
    """results.py:784-853: 'T' then the unknowns; no row for tstart."""
    sol_type = "TRAN"

    def __init__(self, circ, raw, tstart, tstop, method, engine=None):
        solution.__init__(self, circ, dict(raw), engine)
        names, units = _varnames(circ, upper=True)
        self.variables = ["T"] + names
        self.units = dict(units)
        self.units["T"] = "s"
        self.tstart, self.tstop, self.method = tstart, tstop, method

    # With an engine behind it the solution reads PACKED rows: the kernel writes ragged
    # [B][max_rows] buffers (rows[b] of them used: C4 ~1790 of 3072), `Engine.pack_tran` packs
    # the rows that exist contiguously on the device and ONE device->host copy brings them over;
    # the padded device buffers are released.  Per-instance access slices the packed arrays;
    # per-variable access builds a NaN-padded (B, max(rows)) array on the host.
    def _fetch_packed(self):
        if self._np.get('_packed_rows') is not None or self._engine is None:
            return self._np.get('_packed_rows')
        if not (hasattr(self.raw.get('t'), 'is_cuda') and self.raw['t'].is_cuda):
            return None
        pk = self._engine.pack_tran(self.raw)
        host = self._engine.to_host_owned(pk)
        rows = host['rows']
        total = int(rows.sum())
        offs = np.zeros(len(rows) + 1, np.int64)
        np.cumsum(rows, out=offs[1:])
        P = dict(t=host['t'][:total], x=host['x'][:total], offsets=offs)
        self._np.update(rows=rows, counters=host['counters'], status=host['status'], _packed_rows=P)
        self._shape = tuple(self.raw['t'].shape)
        self.raw = {k: (None if k in ('t', 'x', '_packed') else v) for k, v in self.raw.items()}
        return P

    def _get(self, key):
        if key in ('t', 'x') and key not in self._np:
            P = self._fetch_packed()
            if P is not None:                     # padded view of the packed rows, built on demand
                rows, offs = self._np['rows'], P['offsets']
                width = int(rows.max()) if len(rows) else 0
                shape = (len(rows), width) + P[key].shape[1:]
                a = np.full(shape, np.nan)
                for b in range(len(rows)):
                    a[b, :rows[b]] = P[key][offs[b]:offs[b + 1]]
                self._np[key] = a
                return a
        elif key in ('rows', 'counters', 'status') and key not in self._np:
            self._fetch_packed()
        return solution._get(self, key)

    @property
    def rows(self):
        return self._get('rows')

    @property
    def counters(self):
        """(B, 4): dc_solve calls, NR iterations, LTE rejections, NR step cuts."""
        return self._get('counters')

    def _masked(self, a):
        rows = self.rows
        m = np.arange(a.shape[1])[None, :] >= rows[:, None]
        a = np.array(a, copy=True)
        a[m] = np.nan
        return a

    def __getitem__(self, name):
        i = self._index(name)
        P = self._fetch_packed()
        if P is not None:
            rows, offs = self._np['rows'], P['offsets']
            col = P['t'] if i == 0 else P['x'][:, i - 1]
            a = np.full((len(rows), int(rows.max()) if len(rows) else 0), np.nan)
            m = np.arange(a.shape[1])[None, :] < rows[:, None]
            a[m] = col                           # row-major fill == instance-major packed order
            return a
        if i == 0:
            return self._masked(self._get('t'))
        return self._masked(self._get('x')[:, :, i - 1])

    def instance_array(self, b):
        return self.instance(b)

    def instance(self, b):
        """(1 + n, rows[b]) array of instance b — the reference's asarray()."""
        P = self._fetch_packed()
        if P is not None:
            lo, hi = P['offsets'][b], P['offsets'][b + 1]
            return np.concatenate((P['t'][None, lo:hi], P['x'][lo:hi].T), axis=0)
        r = int(self.rows[b])
        return np.concatenate((self._get('t')[b, None, :r],
                               self._get('x')[b, :r, :].T), axis=0)

    def packed(self):
        """(t [R], x [R, n], offsets [B + 1]): every accepted row of every instance, instance
        after instance — the cheapest host view of a large transient batch."""
        P = self._fetch_packed()
        if P is None:
            rows = self.rows
            offs = np.zeros(len(rows) + 1, np.int64)
            np.cumsum(rows, out=offs[1:])
            t, x = self._get('t'), self._get('x')
            return (np.concatenate([t[b, :rows[b]] for b in range(len(rows))]),
                    np.concatenate([x[b, :rows[b]] for b in range(len(rows))]), offs)
        return P['t'], P['x'], P['offsets']

    def get_x(self):
        return self['T']

    def get_xlabel(self):
        return "T"


class ac_solution(solution):
    """results.py:554-712: 'f' [Hz] then complex unknowns."""
    sol_type = "AC"

    def __init__(self, circ, raw, stype, engine=None):
        solution.__init__(self, circ, raw, engine)
        names, units = _varnames(circ, upper=False)
        self.variables = ["f"] + names
        self.units = dict(units)
        self.units["f"] = "Hz"
        self.stype = stype

    def __getitem__(self, name):
        i = self._index(name)
        if i == 0:
            return self.raw['omegas'] / np.pi / 2
        return self._get('x')[:, :, i - 1]

    def asarray(self):
        x = self._get('x')
        f = np.broadcast_to((self.raw['omegas'] / np.pi / 2)[None, None, :],
                            (x.shape[0], 1, x.shape[1])).astype(np.complex128)
        return np.concatenate((f, np.transpose(x, (0, 2, 1))), axis=1)

    def get_x(self):
        return self['f']

    def get_xlabel(self):
        return "f"
