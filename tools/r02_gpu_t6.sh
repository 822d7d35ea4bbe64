#!/bin/bash
cd /root/repo
echo "pairs"; timeout 200 python tools/res_time.py 1024x1024 1024x512 400x400 1024x880 1024x1332 2>&1 | tail -5 | cut -c1-60
echo "no pairs"; PCD_RES_NO_PAIRS=1 timeout 200 python tools/res_time.py 1024x1024 1024x512 400x400 1024x880 1024x1332 2>&1 | tail -5 | cut -c1-60
echo "cluster 8"; PCD_RES_CLUSTER=8 timeout 200 python tools/res_time.py 1024x1024 1024x512 400x400 2>&1 | tail -3 | cut -c1-60
