"""Hand-written sm_100a operators (ctypes bindings over csrc/)."""
