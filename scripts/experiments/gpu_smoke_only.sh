#!/bin/bash
timeout 70 python __graft_entry__.py smoke 2>&1 | tail -1
