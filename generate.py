#!/usr/bin/env python
"""`python generate.py --model runs/kinetic-gan/exp1/models/generator_N.pth ...` - the reference's generation script name,
options and output files (generate.py:27-47,105-123); see kinetic-gan_b200/generate.py."""
import kgan_b200  # noqa: F401  (registers the importable alias of kinetic-gan_b200/)
from kgan_b200.generate import main

if __name__ == "__main__":
    main()
