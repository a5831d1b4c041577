// TEST INFRASTRUCTURE (oracle/): RAJA is included but never used by the reference Functional kernels.
