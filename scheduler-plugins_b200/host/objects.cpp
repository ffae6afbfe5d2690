// Host-side rules of the reference, restated for the C++ host layer (see objects.hpp).
#include "objects.hpp"

#include <algorithm>
#include <set>

namespace b200host {

// apimachinery resource.Quantity [upstream]: <number><suffix>, binary (Ki..Ei), decimal (n u m k M G T P E)
// or decimal exponent (e3).  Exact rational arithmetic in __int128 milli-units.
int64_t ParseQuantity(const std::string& in) {
  std::string s = in;
  while (!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
  size_t i = 0;
  while (i < s.size() && isspace((unsigned char)s[i])) ++i;
  bool neg = false;
  if (i < s.size() && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
  __int128 mant = 0;
  int frac_digits = 0;
  bool seen_digit = false, seen_dot = false;
  for (; i < s.size(); ++i) {
    if (isdigit((unsigned char)s[i])) {
      mant = mant * 10 + (s[i] - '0');
      if (seen_dot) ++frac_digits;
      seen_digit = true;
    } else if (s[i] == '.' && !seen_dot) {
      seen_dot = true;
    } else {
      break;
    }
  }
  if (!seen_digit) throw std::invalid_argument("bad quantity: " + in);
  std::string suf = s.substr(i);
  __int128 num = mant * 1000, den = 1;  // milli-units
  for (int k = 0; k < frac_digits; ++k) den *= 10;
  auto pow10 = [](int e) {
    __int128 r = 1;
    for (int k = 0; k < e; ++k) r *= 10;
    return r;
  };
  if (suf == "Ki") num <<= 10;
  else if (suf == "Mi") num <<= 20;
  else if (suf == "Gi") num <<= 30;
  else if (suf == "Ti") num <<= 40;
  else if (suf == "Pi") num <<= 50;
  else if (suf == "Ei") num <<= 60;
  else if (suf == "n") den *= pow10(9);
  else if (suf == "u") den *= pow10(6);
  else if (suf == "m") den *= 1000;
  else if (suf == "") {}
  else if (suf == "k") num *= pow10(3);
  else if (suf == "M") num *= pow10(6);
  else if (suf == "G") num *= pow10(9);
  else if (suf == "T") num *= pow10(12);
  else if (suf == "P") num *= pow10(15);
  else if (suf == "E") num *= pow10(18);
  else if ((suf[0] == 'e' || suf[0] == 'E') && suf.size() > 1) {
    int e = std::stoi(suf.substr(1));
    if (e >= 0) num *= pow10(e); else den *= pow10(-e);
  } else {
    throw std::invalid_argument("bad quantity suffix: " + in);
  }
  if (num % den != 0) throw std::invalid_argument("quantity finer than one milli-unit (unsupported): " + in);
  __int128 v = num / den;
  if (v > (__int128)INT64_MAX) throw std::invalid_argument("quantity overflows int64 milli-units: " + in);
  return neg ? -(int64_t)v : (int64_t)v;
}

bool IsNativeResource(const std::string& n) {
  return n.find('/') == std::string::npos || n.find("kubernetes.io/") != std::string::npos;
}
bool IsHugePageResourceName(const std::string& n) { return n.rfind("hugepages-", 0) == 0; }
bool IsNUMAAffineResource(const std::string& n) {
  return n == ResourceCPU || n == ResourceMemory || IsHugePageResourceName(n);
}
bool IsHostLevelResource(const std::string& n) {
  return n == ResourceEphemeralStorage || n == "storage" || !IsNativeResource(n);
}
bool IsScalarResourceName(const std::string& n) {
  bool extended = !IsNativeResource(n) && n.rfind("requests.", 0) != 0;
  return extended || IsHugePageResourceName(n) || n.find("kubernetes.io/") != std::string::npos ||
         n.rfind("attachable-volumes-", 0) == 0;
}

// v1qos.GetPodQOS [upstream]: only cpu and memory count; zero quantities are ignored.
QOS GetPodQOS(const Pod& p) {
  ResourceList requests, limits;
  bool guaranteed = true;
  auto visit = [&](const Container& c) {
    for (auto& [name, q] : c.requests)
      if ((name == ResourceCPU || name == ResourceMemory) && q > 0) requests[name] += q;
    std::set<std::string> found;
    for (auto& [name, q] : c.limits)
      if ((name == ResourceCPU || name == ResourceMemory) && q > 0) {
        found.insert(name);
        limits[name] += q;
      }
    if (found.size() != 2) guaranteed = false;
  };
  for (auto& c : p.containers) visit(c);
  for (auto& c : p.init_containers) visit(c);
  if (requests.empty() && limits.empty()) return QOS::BestEffort;
  if (guaranteed)
    for (auto& [name, r] : requests) {
      auto it = limits.find(name);
      if (it == limits.end() || it->second != r) {
        guaranteed = false;
        break;
      }
    }
  if (guaranteed && requests.size() == limits.size()) return QOS::Guaranteed;
  return QOS::Burstable;
}

bool IncludeNonNative(const Pod& p) {
  for (auto& c : p.init_containers)
    for (auto& kv : c.requests)
      if (!IsNativeResource(kv.first)) return true;
  for (auto& c : p.containers)
    for (auto& kv : c.requests)
      if (!IsNativeResource(kv.first)) return true;
  return false;
}

ResourceList GetPodEffectiveRequest(const Pod& p) {
  ResourceList init, res;
  for (auto& c : p.init_containers)
    for (auto& [name, q] : c.requests) {
      auto it = init.find(name);
      if (it != init.end() && q <= it->second) continue;
      init[name] = q;
    }
  for (auto& c : p.containers)
    for (auto& [name, q] : c.requests) res[name] += q;  // first occurrence inserts q
  for (auto& [name, q] : init) {
    auto it = res.find(name);
    if (it != res.end() && q <= it->second) continue;
    res[name] = q;
  }
  if (p.has_overhead)
    for (auto& [name, q] : p.overhead) res[name] += q;
  return res;
}

static double GoRound(double x) {  // math.Round: half away from zero
  double a = std::fabs(x), f = std::floor(a);
  if (a - f >= 0.5) f += 1.0;
  return std::copysign(f, x);
}

int64_t PredictUtilisation(const Container& c, int64_t default_milli, double multiplier) {
  auto l = c.limits.find(ResourceCPU);
  if (l != c.limits.end()) return l->second;
  auto r = c.requests.find(ResourceCPU);
  if (r != c.requests.end()) return (int64_t)GoRound((double)r->second * multiplier);
  return default_milli;
}

int64_t PodPredictedCPU(const Pod& p, int64_t default_milli, double multiplier) {
  int64_t total = 0;
  for (auto& c : p.containers) total += PredictUtilisation(c, default_milli, multiplier);
  if (p.has_overhead) {
    auto it = p.overhead.find(ResourceCPU);
    if (it != p.overhead.end()) total += it->second;
  }
  return total;
}

// GetEffectiveResource: resourcestats.go:124-146 (framework.Resource.Add takes MilliValue for cpu, Value for memory)
static void EffectiveResource(const Pod& p, bool limits, int64_t* cpu_milli, int64_t* mem_bytes) {
  int64_t cpu = 0, mem = 0;
  auto get = [](const ResourceList& r, const char* k, int64_t* out) {
    auto it = r.find(k);
    if (it == r.end()) return false;
    *out = it->second;
    return true;
  };
  int64_t v;
  for (auto& c : p.containers) {
    const ResourceList& rl = limits ? c.limits : c.requests;
    if (get(rl, ResourceCPU, &v)) cpu += v;
    if (get(rl, ResourceMemory, &v)) mem += QuantityValue(v);
  }
  for (auto& c : p.init_containers) {
    const ResourceList& rl = limits ? c.limits : c.requests;
    if (get(rl, ResourceCPU, &v)) cpu = std::max(cpu, v);
    if (get(rl, ResourceMemory, &v)) mem = std::max(mem, QuantityValue(v));
  }
  if (p.has_overhead) {
    if (get(p.overhead, ResourceCPU, &v)) cpu += v;
    if (get(p.overhead, ResourceMemory, &v)) mem += QuantityValue(v);
  }
  *cpu_milli = cpu;
  *mem_bytes = mem;
}
void GetResourceRequested(const Pod& p, int64_t* cpu_milli, int64_t* mem_bytes) { EffectiveResource(p, false, cpu_milli, mem_bytes); }
void GetResourceLimits(const Pod& p, int64_t* cpu_milli, int64_t* mem_bytes) { EffectiveResource(p, true, cpu_milli, mem_bytes); }

void NodeRequestsAndLimitsOfRunningPods(const NodeInfo& ni, int64_t out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0;
  for (auto& p : ni.pods) {
    if (!p) continue;
    int64_t qc, qm, xc, xm;
    GetResourceRequested(*p, &qc, &qm);
    GetResourceLimits(*p, &xc, &xm);
    xc = std::max(xc, qc), xm = std::max(xm, qm);  // SetMaxLimits :230-246
    out[0] += qc, out[1] += qm, out[2] += xc, out[3] += xm;
  }
}

int64_t GetResourceRequestQuantityCPU(const Pod& p) {
  int64_t total = 0;
  for (auto& c : p.containers) {
    auto it = c.requests.find(ResourceCPU);
    if (it != c.requests.end()) total += it->second;
  }
  for (auto& c : p.init_containers) {
    auto it = c.requests.find(ResourceCPU);
    if (it != c.requests.end() && total < it->second) total = it->second;
  }
  if (p.has_overhead) {
    auto it = p.overhead.find(ResourceCPU);
    if (it != p.overhead.end() && total != 0) total += it->second;
  }
  return total;
}

void GetResourceData(const std::vector<Metric>& ms, const std::string& type, double* avg, double* std_, bool* valid) {
  *avg = 0;
  *std_ = 0;
  *valid = false;
  bool avg_found = false;
  for (auto& m : ms) {
    if (m.type != type) continue;
    if (m.op == "AVG") {
      *avg = m.value;
      avg_found = true;
    } else if (m.op == "STD") {
      *std_ = m.value;
    } else if ((m.op.empty() || m.op == "Latest") && !avg_found) {
      *avg = m.value;
    }
    *valid = true;
  }
}

TopologyManager TopologyManagerFromNodeResourceTopology(const NodeResourceTopology& nrt) {
  TopologyManager c;
  if (!nrt.topology_policies.empty()) {  // updateFromPolicies: only the first entry
    const std::string& p = nrt.topology_policies[0];
    static const std::map<std::string, std::pair<const char*, const char*>> table = {
        {"SingleNUMANodePodLevel", {"single-numa-node", "pod"}},
        {"SingleNUMANodeContainerLevel", {"single-numa-node", "container"}},
        {"BestEffortPodLevel", {"best-effort", "pod"}},
        {"BestEffortContainerLevel", {"best-effort", "container"}},
        {"RestrictedPodLevel", {"restricted", "pod"}},
        {"RestrictedContainerLevel", {"restricted", "container"}}};
    auto it = table.find(p);
    if (it != table.end()) {
      c.policy = it->second.first;
      c.scope = it->second.second;
    }
  }
  for (auto& [name, value] : nrt.attributes) {  // updateFromAttributes
    if (name == "topologyManagerScope" && (value == "container" || value == "pod")) c.scope = value;
    else if (name == "topologyManagerPolicy" &&
             (value == "none" || value == "best-effort" || value == "restricted" || value == "single-numa-node"))
      c.policy = value;
    else if (name == "topologyManagerMaxNUMANodes") {
      try {
        size_t pos = 0;
        int v = std::stoi(value, &pos);
        if (pos == value.size() && v > 1) c.max_numa_nodes = std::min(v, 1024);
      } catch (...) {
      }
    }
  }
  return c;
}

}  // namespace b200host
