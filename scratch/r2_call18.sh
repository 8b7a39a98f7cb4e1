#!/bin/bash
cd /root/repo
timeout 900 python scratch/debug_rim.py MATE-4v8-9.yaml 65536 64457 6 1 45 2>&1 | tail -14
timeout 900 python scratch/debug_rim.py MATE-Navigation.yaml 65536 31773 3 12 20 2>&1 | tail -14
