"""Plain-torch stand-in for the ~10 torch_geometric 2.0.4 symbols the reference's encoder path imports
(dependency.txt:2 pins torch_geometric==2.0.4; the package is not installable here — SURVEY.md 8c).

TEST INFRASTRUCTURE ONLY (lives under oracle/).  With this directory on sys.path the reference's
model/gnn.py and model/model.py import and run UNMODIFIED; tests/golden/gen_encoder_golden.py uses that to
produce encoder golden vectors.  The semantics restated here are recalled from the PyG 2.0.4 sources
(they cannot be re-read in this container) and are listed in oracle/README.md; the encoder oracle is
therefore "pinned to the reference's own model code, on a recalled PyG layer".
"""
__version__ = "2.0.4-shim"
from . import data, loader, nn, transforms  # noqa: F401
