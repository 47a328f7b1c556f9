#!/bin/bash
timeout 600 python tools/dbg_gradnorm_test.py 2>&1 | tail -12
