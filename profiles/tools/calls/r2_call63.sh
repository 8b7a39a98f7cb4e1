#!/bin/bash
cd /root/repo
AB_ARGS=--no-configs bash scratch/ab.sh keep keep@MATE_B200_L2_KEEP=16 keep@MATE_B200_L2_KEEP=64
