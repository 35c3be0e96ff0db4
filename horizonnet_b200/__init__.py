"""horizonnet_b200 -- B200-native (sm_100a) implementation of HorizonNet's inference hot path.

Public surface = the reference's own:
    from horizonnet_b200.model import HorizonNet                 # reference model.py
    from horizonnet_b200.misc.panostretch import pano_stretch    # reference misc/panostretch.py
"""
__all__ = ['model', 'misc', 'weights']
