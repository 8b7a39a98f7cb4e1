#!/bin/bash
cd /root/repo
AB_ARGS=--no-configs bash scratch/ab.sh u1 u2 obfixed
