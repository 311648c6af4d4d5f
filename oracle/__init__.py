"""TEST INFRASTRUCTURE -- CPU oracle for the C2A CCD hot path (see oracle/c2a_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (c2a_b200) never does.
"""
from .binding import (  # noqa: F401
    OrcResult, RESULT_DTYPE, CONTACT_DTYPE, DISTANCE_DTYPE, build_oracle, have_ref, port, ref, RefModel, bvh_struct,
)
