#!/usr/bin/env python3
"""Write a docker-compose file for a multi-host elastic job on one machine: a config-server container, a client container that
keeps PUTting randomly resized clusters, and N node containers each running `kungfu-run -w -k` on its own IP of a private subnet.

Reference: benchmarks/adaptation/gen-compose.py:60-121 and tests/go/cmd/kungfu-config-client-example (every second a random-size,
shuffled subset). Without docker the same tier runs on loopback aliases (tests/test_multihost_loopback.py).

    python benchmarks/adaptation/gen_compose.py --nodes 4 --slots 2 > docker-compose.yaml && docker compose up
"""
import argparse


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=4)
    ap.add_argument("--slots", type=int, default=2)
    ap.add_argument("--image", default="kungfu-b200:cpu")
    ap.add_argument("--subnet", default="172.28.0")
    ap.add_argument("--period", type=float, default=2.0)
    ap.add_argument("--steps", type=int, default=60)
    a = ap.parse_args()
    ips = [f"{a.subnet}.{10 + i}" for i in range(a.nodes)]
    hosts = ",".join(f"{ip}:{a.slots}" for ip in ips)
    cs = f"{a.subnet}.2"
    url = f"http://{cs}:9100/config"
    out = ["services:",
           "  config-server:", f"    image: {a.image}", "    entrypoint: [\"kungfu-config-server\", \"-port\", \"9100\"]",
           "    networks:", "      kf:", f"        ipv4_address: {cs}",
           "  resizer:", f"    image: {a.image}",
           f"    entrypoint: [\"python\", \"-m\", \"kungfu_b200.cmd.config_client\", \"--url\", \"{url}\", \"--random\", \"--period\", \"{a.period}\","
           f" \"--max\", \"{a.nodes * a.slots}\", \"--duration\", \"{a.steps * 2}\", \"--final\", \"{a.slots}\"]",
           "    depends_on: [config-server]", "    networks:", "      kf:", f"        ipv4_address: {a.subnet}.3"]
    for i, ip in enumerate(ips):
        out += [f"  node{i}:", f"    image: {a.image}",
                f"    entrypoint: [\"kungfu-run\", \"-w\", \"-k\", \"-nic\", \"eth0\", \"-np\", \"{a.slots}\", \"-H\", \"{hosts}\", \"-config-server\", \"{url}\","
                f" \"python\", \"benchmarks/adaptation/elastic_resize.py\", \"--steps\", \"{a.steps}\", \"--follow-config-server\"]",
                "    depends_on: [config-server]", "    networks:", "      kf:", f"        ipv4_address: {ip}"]
    out += ["networks:", "  kf:", "    ipam:", "      config:", f"        - subnet: {a.subnet}.0/24"]
    print("\n".join(out))


if __name__ == "__main__":
    main()
