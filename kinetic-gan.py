#!/usr/bin/env python
"""`python kinetic-gan.py --data_path ... --label_path ...` - the reference's training script name and options
(kinetic-gan.py:23-44), running on the B200 trainer; see kinetic-gan_b200/train.py.
DDP: `python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 kinetic-gan.py ...`."""
import kgan_b200  # noqa: F401  (registers the importable alias of kinetic-gan_b200/)
from kgan_b200.train import main

if __name__ == "__main__":
    main()
