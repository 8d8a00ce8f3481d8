#!/bin/bash
# round 2, GPU call M: the whole GPU suite after the register tridiagonalisation and the odd-cohort alignment fix of the permutation kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r2m_pytest.log
