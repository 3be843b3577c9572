"""Copy the world-layout DATA files (map_txt/*.txt: ';'-separated rows of W/S/@/space symbols) from the
reference tree into the package so `env_layout_file=` keeps working.  Data fixtures, not source code."""
import glob
import os
import shutil

SRC = "/root/reference/ai_economist/foundation/scenarios/simple_wood_and_stone/map_txt"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ai_economist_b200", "foundation", "map_txt")
if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for f in glob.glob(SRC + "/*.txt"):
        shutil.copy(f, DST)
        print("copied", os.path.basename(f))
