"""Run the reference's UNMODIFIED entry script against the B200 classes.

    python -m anomalyclip_b200.dropin /path/to/AnomalyCLIP/src/eval.py data=ucfcrime \
        model=anomaly_clip_ucfcrime ckpt_path=/path/to/last.ckpt

`src/eval.py` puts its own project root at the front of sys.path (pyrootutils.setup_root,
src/eval.py:9) and then does `from src import utils` (:27), so started directly it would import the
reference's whole `src` package.  This launcher imports THIS repository's `src` shim first (with the
reference's `src/` directory appended to its search path, see src/__init__.py): the later
`from src import utils` finds the already-imported package, `src.utils` comes from the reference,
and the `_target_` strings of the configs (`src.models.anomaly_clip_module.AnomalyCLIPModule`,
`src.models.components.anomaly_clip.AnomalyCLIP`, `src.data.anomaly_clip_datamodule.
AnomalyCLIPDataModule`) resolve to the B200 implementations.  Hydra, Lightning and the reference's
configs are used as they are."""
from __future__ import annotations

import os
import runpy
import sys
from pathlib import Path


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = Path(argv[0]).resolve()
    if not script.is_file():
        raise SystemExit(f"dropin: {script} not found")
    ref_root = script.parent.parent                      # <root>/src/eval.py
    os.environ.setdefault("ACLIP_REFERENCE_ROOT", str(ref_root))
    repo_root = str(Path(__file__).resolve().parent.parent)
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    import src                                            # this repository's shim, before the script's
    here = Path(src.__file__).resolve().parent
    if here.parent != Path(repo_root):
        raise SystemExit(f"dropin: 'src' resolved to {here}, not to this repository's shim")
    import src.data.anomaly_clip_datamodule  # noqa: F401  (fail early if the shim is incomplete)
    import src.models.anomaly_clip_module  # noqa: F401
    sys.argv = [str(script)] + argv[1:]
    runpy.run_path(str(script), run_name="__main__")


if __name__ == "__main__":
    main()
