"""CPU oracle for the quadcopter env-step hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (torch-CPU / numpy, float32 or float64) of the reference
algorithm that `ouzelum_b200`'s CUDA kernels implement.  Every function cites the reference
file:line it follows (paths relative to the reference checkout's root).

Rules (see DESIGN.md "Oracle"):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
    legs may import anything from here -- as the checker or as the timed CPU baseline, never as
    the product.  `ouzelum_b200/` never imports `oracle` and has no CPU fallback.
  * nothing here reads /root/reference at run time; the pins against the reference's own code
    were generated once by `tests/golden/make_golden.py` (committed) into `tests/golden/*.npz`.

Parity status: the restated torch functions (reward, observation, Lee controllers, PV filter,
waypoint tables, differential drive) are PINNED against the reference's own code executed in the
build container (fixtures in tests/golden/).  The rigid-body integrator (PhysX, closed source, not
under the reference tree) and the AHRS-EKF's third-party `ahrs` helpers are "parity unpinned":
restated from the published algorithm / SURVEY.md section 8a row P, see DESIGN.md.
"""
