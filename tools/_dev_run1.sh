python examples/specialize_robot.py 2>&1 | tail -5
