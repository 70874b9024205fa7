#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|rror|FAILED|rel err|max-abs" | tail -25
