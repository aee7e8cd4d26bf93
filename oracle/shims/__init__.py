"""Stand-ins for the third-party modules the reference's Python files import but this image lacks
(TEST / BASELINE INFRASTRUCTURE, see oracle/ref_gpu_vo.py).  Nothing under rampvo_b200/ imports them."""
