"""Import shim for the reference's ``from data_utils import *`` -- see compat/mtn.py."""
from mtn_b200.data_utils import *     # noqa: F401,F403
from mtn_b200.data_utils import (Batch, NoamOpt, SimpleLossCompute, subsequent_mask, encode, greedy_decode,  # noqa: F401
                                 beam_search_decode)
