"""Drop-in tree: put this directory first on sys.path and the reference scripts' own imports
(`from models.unet import UNet`, `from models.consistency_reglur_memory import ConsistencyRegulr`,
`from losses.losses import *`; train_chaos.py:17,20, upper_bound_chaos.py:18,21, inference.py:22,25)
resolve to the B200-native implementations."""
import os

DROPIN_PATH = os.path.dirname(os.path.abspath(__file__))
