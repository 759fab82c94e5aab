"""B200-native NanoWrap conjugate-gradient shrinkwrap hot path (drop-in for ch_shrinkwrap's
``ShrinkwrapMeshConjGrad`` / ``MembraneMesh.shrink_wrap`` / ``ShrinkwrapMembrane``)."""
__version__ = '0.1.0'
