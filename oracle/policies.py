"""Scripted action sequences for parity runs -- TEST INFRASTRUCTURE.

Random actions almost never contain the fire, so the containment branch of
World.get_reward (environment.py:345-377) would go untested.  ``ring_actions``
walks the bulldozer once round a square ring about the fire origin, which
encloses the fire before its first spread at tick 19.
Actions: 0 N (y-1), 1 S (y+1), 2 E (x+1), 3 W (x-1)  (environment.py:163-171).
"""
from __future__ import annotations


def ring_actions(ax: int, ay: int, cx: int, cy: int, R: int):
    """Actions taking the agent from (ax, ay) onto the Chebyshev-radius-R ring round
    (cx, cy) and once clockwise along it (plus 2 steps of overlap)."""
    acts = []
    x, y = ax, ay

    def go(a):
        nonlocal x, y
        acts.append(a)
        dx, dy = [(0, -1), (0, 1), (1, 0), (-1, 0)][a]
        x, y = x + dx, y + dy

    # 1. step outwards until on the ring
    while max(abs(x - cx), abs(y - cy)) < R:
        if abs(x - cx) >= abs(y - cy):
            go(2 if x >= cx else 3)
        else:
            go(1 if y >= cy else 0)
    # 2. walk the ring clockwise (screen coords, y down): top edge -> E, right edge -> S, ...
    for _ in range(8 * R + 2):
        dx, dy = x - cx, y - cy
        if dy == -R and dx < R:
            go(2)
        elif dx == R and dy < R:
            go(1)
        elif dy == R and dx > -R:
            go(3)
        else:
            go(0)
    return acts
