#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest "$@" -x -q 2>&1 | tail -30
