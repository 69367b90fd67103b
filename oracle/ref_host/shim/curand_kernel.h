// TEST INFRASTRUCTURE ONLY (oracle/): empty stand-in; cuda_shim.h (force-included) supplies everything.
