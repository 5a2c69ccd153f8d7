#!/bin/bash
timeout 300 python tools/conv_phases.py 2>&1 | tail -5
