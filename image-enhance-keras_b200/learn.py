"""Drop-in mirror of the reference learn.py (learn.py:1-22): train DifvdsrDouble for 180 epochs."""
from __future__ import print_function, division

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import models  # noqa: E402

if __name__ == "__main__":
    scale = 1
    ddsr = models.DifvdsrDouble(scale)
    ddsr.create_model()
    ddsr.fit(nb_epochs=180)
