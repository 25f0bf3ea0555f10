#!/bin/bash
# compute-sanitizer memcheck over the small-size parity tests of every kernel family (run on the GPU box)
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x \
  -k "not 3840 and not 1920 and not size4 and not size5 and not size3 and not 4k and not pool and not sdk" 2>&1 | tail -8
