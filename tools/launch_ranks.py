#!/usr/bin/env python
"""One process per GPU for a plain executable (what torchrun does for Python scripts, mpirun for MPI
binaries): sets RANK / WORLD_SIZE / LOCAL_RANK / CFD2D_JOB_DIR / CFD2D_JOB_ID and waits.
   python tools/launch_ranks.py N cfd-2d_b200/host/_build/cfd2d_cuda task.xml"""
import os
import subprocess
import sys
import uuid


def main():
    n = int(sys.argv[1])
    cmd = sys.argv[2:]
    job = uuid.uuid4().hex[:12]
    procs = []
    for r in range(n):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(n), LOCAL_RANK=str(r), CFD2D_JOB_DIR=os.getcwd(), CFD2D_JOB_ID=job)
        procs.append(subprocess.Popen(cmd, env=env))
    rc = 0
    for p in procs:
        rc = max(rc, abs(p.wait()))
    sys.exit(rc)


if __name__ == "__main__":
    main()
