"""Frame loops around ``Scene.render`` (reference: sightpy/animation.py:6-54)."""
from pathlib import Path

import numpy as np

__all__ = ["create_animation", "create_animation_using_opencv"]


def _frame_times(fps, start_time, final_time):
    count = int(fps * (final_time - start_time))
    dt = (final_time - start_time) / count
    return [(i, start_time + i * dt) for i in range(count)]


def create_animation(scene, samples_per_pixel, fps, start_time, final_time, update_scene, name):
    """Render frames to ./frames/<name>_<i>.png (assemble with ffmpeg)."""
    Path("./frames").mkdir(exist_ok=True)
    for i, t in _frame_times(fps, start_time, final_time):
        update_scene(scene, t)
        scene.invalidate()
        scene.render(samples_per_pixel).save("frames/" + name + "_" + str(i) + ".png")


def create_animation_using_opencv(scene, samples_per_pixel, fps, start_time, final_time, update_scene, name):
    import cv2
    dims = (scene.camera.screen_width, scene.camera.screen_height)
    video = cv2.VideoWriter(name, cv2.VideoWriter_fourcc("M", "J", "P", "G"), fps, dims)
    for _, t in _frame_times(fps, start_time, final_time):
        update_scene(scene, t)
        scene.invalidate()
        video.write(cv2.cvtColor(np.array(scene.render(samples_per_pixel)), cv2.COLOR_RGB2BGR))
    video.release()
