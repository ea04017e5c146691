#!/bin/bash
timeout 100 python tools/kernel_breakdown.py --cfg 3 4b 5 2>/dev/null | cut -d'|' -f2-9 | tail -3
