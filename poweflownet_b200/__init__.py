"""poweflownet_b200 -- B200-native (sm_100a) implementation of PowerFlowNet's data-parallel hot path:
`MaskEmbdMultiMPN` forward + backward (reference networks/MPN.py:6-56,456-559) behind the reference's
own `nn.Module` surface.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
