cd $GRAFT_REPO_ROOT
timeout 600 python tools/part_probe.py c4s 1 0 CSXB_VALUE_POLICY 0,1 2>&1 | grep GB/s
timeout 600 python tools/part_probe.py c4ns 1 0 CSXB_VALUE_POLICY 0,1 2>&1 | grep GB/s
timeout 600 python tools/part_probe.py c3bs 1 0 CSXB_VALUE_POLICY 0,1 2>&1 | grep GB/s
timeout 600 python tools/part_probe.py c5s 1 0 CSXB_VALUE_POLICY 0,1 2>&1 | grep GB/s
