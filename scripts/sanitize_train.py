"""Training-path subset of sanitize_target.py (compute-sanitizer under gpurun)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sanitize_target.py')).read()
head = src[:src.index('for B, N in [(300, 3)')]
tail = src[src.index('# round 2, training step'):]
exec(head + tail)
