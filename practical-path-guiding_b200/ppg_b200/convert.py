"""Scene converter for C / C++ hosts:  python -m ppg_b200.convert scene.xml scene.ppgscene [width height]

Writes the flat binary form of a Mitsuba 0.5 scene (SceneDesc.save_flat) that ppg_scene_file_load (include/ppg.h) reads; used by the
Mitsuba plugin shim integration/guided_path_b200.cpp, because Mitsuba 0.5 offers no public accessors for nested BSDFs / textures of a live
Scene (twosided, mask, bumpmap keep their children in protected members), while its XML says everything."""
import sys

from .scene import SceneDesc, load_mitsuba_xml


def main(argv):
    if len(argv) < 3:
        print(__doc__); return 2
    sc = SceneDesc.load(argv[1]) if argv[1].endswith(".npz") else load_mitsuba_xml(argv[1])
    if len(argv) >= 5:
        sc = sc.with_film(int(argv[3]), int(argv[4]))
    sc.save_flat(argv[2])
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
