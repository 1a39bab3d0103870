"""TEST INFRASTRUCTURE: CPU oracle for the SWIPDG hot path (see swipdg_oracle.cpp)."""
