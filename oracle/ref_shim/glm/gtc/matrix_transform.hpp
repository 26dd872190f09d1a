// intentionally empty (oracle build only)
#pragma once
