import json, sys
d = json.load(open(sys.argv[1]))
print("MDE/s", round(d["value"]), "fps", round(d["fps"], 2), "e2e fps", round(d["e2e"]["fps"], 2), "launches", d["gpu_launches"])
for k, v in d["stages"].items():
    print(" ", k, v)
print(" clocks", d["clocks"], "bad2", d["quality"])
