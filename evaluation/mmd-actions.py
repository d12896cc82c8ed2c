#!/usr/bin/env python
"""`python evaluation/mmd-actions.py --data_real ... --labels_real ... --data_fake ... --labels_fake ...` - the reference's MMD
script name and options (evaluation/mmd-actions.py:120-128); see kinetic-gan_b200/evaluation.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgan_b200  # noqa: E402,F401  (registers the importable alias of kinetic-gan_b200/)
from kgan_b200.evaluation import main  # noqa: E402

if __name__ == "__main__":
    main()
