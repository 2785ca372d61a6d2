"""deep_mcts_b200 -- B200-native batched MCTS self-play engine with the API of henribru/deep-mcts.

The hot path (game engines, PUCT tree search, rollouts, ConvolutionalNet leaf evaluation) runs
as hand-written sm_100a CUDA kernels in ``libdeepmcts_b200.so`` behind the C ABI declared in
``include/deepmcts_b200.h``; this package is the host-side mirror of the reference's
``GameManager`` / ``GameNet`` / ``MCTS`` interface.  There is no CPU fallback.
"""
__version__ = "0.1.0"
