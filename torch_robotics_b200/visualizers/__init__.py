"""Only the two drawing helpers the reference's IK example calls (examples/inverse_kinematics.py:47-84); the planning
visualisers are outside this repository's scope."""
