"""The batch axis: per-instance overrides of element / model parameters.

The reference has no batch mode — a Monte-Carlo or parameter sweep is a Python
loop that builds a fresh `Circuit` per instance (SURVEY.md §8c "oracle loop
caveats").  `ParamBatch` is the batched equivalent of that loop's inputs: for
each varied quantity, a `(B,)` float64 array named after the *constructor
argument* it replaces (e.g. ``('R1', 'value')``, ``('di', 'IS')``,
``('nch', 'VTO')``, ``('M1', 'W')``).  Derived quantities (ekv VTO sign flip,
COX from TOX, ...) are re-derived per instance by `compile.flatten`."""
import numpy as np


class ParamBatch(object):
    def __init__(self, B):
        self.B = int(B)
        self._ov = {}
        self._version = 0        # bumped by set(): compile.flatten caches per (circuit, version)
        self._flat_cache = None

    def set(self, owner, attr, values):
        """owner: element part_id or model label (case-insensitive);
        attr: constructor keyword; values: array-like of shape (B,)."""
        v = np.ascontiguousarray(values)
        if v.shape != (self.B,):
            raise ValueError("batch values for %s.%s must have shape (%d,), "
                             "got %s" % (owner, attr, self.B, v.shape))
        if not np.iscomplexobj(v):
            v = v.astype(np.float64)
        self._ov[(owner.lower(), attr)] = v
        self._version += 1
        self._flat_cache = None
        return self

    def __setitem__(self, key, values):
        self.set(key[0], key[1], values)

    def __getitem__(self, key):
        return self._ov[(key[0].lower(), key[1])]

    def __len__(self):
        return self.B

    def items(self):
        return self._ov.items()

    def for_owner(self, owner):
        o = owner.lower()
        return {a: v for (ow, a), v in self._ov.items() if ow == o}

    def owners(self):
        return set(ow for (ow, _a) in self._ov)

    def slice(self, lo, hi):
        """Instances [lo, hi) — what one rank of a sharded run owns."""
        out = ParamBatch(hi - lo)
        for k, v in self._ov.items():
            out._ov[k] = np.ascontiguousarray(v[lo:hi])
        return out
