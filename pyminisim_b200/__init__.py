"""pyminisim_b200 -- B200-native (sm_100a) pedestrian-dynamics hot path of TimeEscaper/pyminisim.

Batched API : pyminisim_b200.batched.BatchedSimulation / BatchedOrca   (W worlds per GPU)
Drop-in API : pyminisim_b200.core / .pedestrians / .sensors / .world_map / .robot  (same names as pyminisim.*)
C ABI       : include/pms.h, implemented by pyminisim_b200/libpms_b200.so (build: python -m pyminisim_b200.build)
"""
__version__ = "0.1.0"
