"""
Importable name of the package whose sources live in `two-towers-overlords_b200/` (a hyphen cannot appear in
a Python import).  Sub-modules resolve through the extended __path__:

    from two_towers_overlords_b200 import TwoTowersModel, TripletLoss
    from two_towers_overlords_b200.training import run_training, evaluate_model, FusedTrainer

Drop-in mode for the reference's backend/main.py: put `two-towers-overlords_b200/` itself on sys.path, then
`import model`, `import training`, `import data` resolve to these modules.
"""
import os as _os

IMPL_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "two-towers-overlords_b200")
__path__.append(IMPL_DIR)

from . import _native  # noqa: E402  (loads nothing until first use; raises loudly if the .so is missing)
from .model import AveragePoolingTower, TokenBatch, TripletLoss, TwoTowersModel  # noqa: E402,F401
