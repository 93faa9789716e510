"""ahkab_b200 — B200-native batched Newton/MNA engine behind ahkab's API.

    import ahkab_b200 as ahkab
    cir = ahkab.Circuit('...'); cir.add_resistor(...); ...
    batch = ahkab.ParamBatch(B); batch.set('R1', 'value', values)
    res = ahkab.run(cir, ahkab.new_op(), batch=batch)
    res['op']['VN1']        # (B,) array

Importing the package does not need a GPU; running an analysis does (the CUDA
library has no CPU fallback)."""
from . import options, constants, time_functions, netlist   # noqa: F401
from .circuit import Circuit                                  # noqa: F401
from .batch import ParamBatch                                 # noqa: F401
from .analyses import (new_op, new_dc, new_tran, new_ac, queue, run,   # noqa: F401
                       new_x0, icmodified_x0, get_op_x0, get_engine)
from .engine import Engine, EngineError                       # noqa: F401

__version__ = "0.1.0"
