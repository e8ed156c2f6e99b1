"""Zero-edit drop-in.  Put this directory BEFORE the reference checkout on PYTHONPATH,

    PYTHONPATH=/path/to/pips-b200/shim:/path/to/pips-b200:/path/to/aharley-pips python demo.py

and ``from nets.pips import Pips`` (demo.py:9, chain_demo.py:9, test_on_*.py, train.py) resolves to this file.  The
reference's ``nets`` directory has no ``__init__.py`` (a namespace package), and neither has this one, so every other
module of the package (nets.raftnet, nets.raft_core, ...) is still found in the reference checkout.

``nets.pips`` of the reference (nets/pips.py:400-611), served by pips_b200: same class name, constructor, forward
signature, return tuples and state_dict keys.  The loss helpers the reference's training
script imports from this module are re-exported from the torch path."""
from pips_b200.pips import DeltaBlock, Pips  # noqa: F401
from pips_b200.torch_path import balanced_ce_loss, score_map_loss, sequence_loss  # noqa: F401
