"""create_fig_and_axes (reference: visualizers/planning_visualizer.py:12-19).  The PlanningVisualizer class is out of scope."""
import matplotlib.pyplot as plt


def create_fig_and_axes(dim=2):
    fig = plt.figure(layout="tight")
    ax = fig.add_subplot(projection="3d") if dim == 3 else fig.add_subplot()
    return fig, ax
