import json, os, sys
json.dump(dict(os.environ), open(sys.argv[1], 'w'), indent=0, sort_keys=True)
