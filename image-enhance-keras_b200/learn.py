"""Drop-in for the reference's training script (learn.py:1-22): `python learn.py` builds DifvdsrDouble(1) and fits it for
180 epochs on the directories img_utils names (learn.py:20-22).  An optional first argument overrides the epoch count
(the reference has none; without it the behaviour is the reference's)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

EPOCHS = 180      # learn.py:22
SCALE = 1         # learn.py:11


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    import models
    net = models.DifvdsrDouble(SCALE)
    net.create_model()
    net.fit(nb_epochs=int(argv[0]) if argv else EPOCHS)
    return net


if __name__ == "__main__":
    main()
