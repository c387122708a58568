cd $GRAFT_REPO_ROOT
timeout 600 python tools/part_probe.py c3 1 0 CSXB_GATHER_POLICY 0,1 2>&1 | grep GB/s
timeout 600 python tools/part_probe.py c3 8 3 CSXB_GATHER_POLICY 0,1 2>&1 | grep GB/s
timeout 600 python tools/part_probe.py c2 1 0 CSXB_GATHER_POLICY 0,1 2>&1 | grep GB/s
