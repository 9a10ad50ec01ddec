"""Import-path shim: the reference's Hydra configs name classes as `src.models...` / `src.data...`
(configs/model/*.yaml:1,16, configs/data/*.yaml:1); the modules below this package re-export the
B200 implementations under those paths.

Everything else of the reference's `src` package (`src.utils`, `src.eval`, `src.train`) stays the
reference's own code: when $ACLIP_REFERENCE_ROOT names a checkout of lucazanella/AnomalyCLIP, its
`src/` directory is appended to this package's search path, so `from src import utils` in the
unmodified `src/eval.py` (src/eval.py:27) finds the reference's `src/utils`, while
`src.models.*` / `src.data.*` resolve here first.  `python -m anomalyclip_b200.dropin` sets this up
and runs the reference's entry script."""
import os as _os

_ref = _os.environ.get("ACLIP_REFERENCE_ROOT")
if _ref and _os.path.isdir(_os.path.join(_ref, "src")):
    __path__.append(_os.path.join(_ref, "src"))
