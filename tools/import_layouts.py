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

# COVID-19 scenario data (fitted parameters, model constants, real-world time series): data, not code
COVID_SRC = "/root/reference/ai_economist/datasets/covid19_datasets/data_and_fitted_params"
COVID_DST = os.path.join(os.path.dirname(DST), "covid19_data")
if __name__ == "__main__":
    os.makedirs(COVID_DST, exist_ok=True)
    for f in ("fitted_params.json", "model_constants.json", "real_world_data.npz"):
        shutil.copy(os.path.join(COVID_SRC, f), COVID_DST)
        print("copied", f)
