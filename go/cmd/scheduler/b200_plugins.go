/*
Copyright The Kubernetes Authors.

Licensed under the Apache License, Version 2.0 (the "License");
you may not use this file except in compliance with the License.
You may obtain a copy of the License at

    http://www.apache.org/licenses/LICENSE-2.0

Unless required by applicable law or agreed to in writing, software
distributed under the License is distributed on an "AS IS" BASIS,
WITHOUT WARRANTIES OR CONDITIONS OF ANY KIND, either express or implied.
See the License for the specific language governing permissions and
limitations under the License.
*/

// The registry delta of the drop-in, kept apart from the reference's main.go so that the change to that file is one
// line: in cmd/scheduler/main.go (reference :50-67) the five app.WithPlugin(...) options of the plugins below are
// deleted and b200Plugins()... is appended to the remaining nine --
//
//	opts := append(b200Plugins(), /* capacityscheduling, coscheduling, topologicalsort, preemptiontoleration,
//	                                 lowriskovercommitment, sysched, peaks, podstate, qos: unchanged */ ...)
//	command := app.NewSchedulerCommand(opts...)
//
// The registry NAMES are the reference's, so an existing KubeSchedulerConfiguration keeps working; every *B200 factory
// degrades to the original plugin when no GPU is usable (b200sched.New fails).  Never compiled in this repository
// (no Go toolchain).
package main

import (
	"k8s.io/kubernetes/cmd/kube-scheduler/app"

	"sigs.k8s.io/scheduler-plugins/pkg/networkaware/networkoverhead"
	"sigs.k8s.io/scheduler-plugins/pkg/noderesources"
	"sigs.k8s.io/scheduler-plugins/pkg/noderesourcetopology"
	"sigs.k8s.io/scheduler-plugins/pkg/trimaran/loadvariationriskbalancing"
	"sigs.k8s.io/scheduler-plugins/pkg/trimaran/targetloadpacking"
)

// b200Plugins: the five data-parallel plugins with the engine-backed factories of their *_b200.go files.
func b200Plugins() []app.Option {
	return []app.Option{
		app.WithPlugin(noderesources.AllocatableName, noderesources.NewAllocatableB200),
		app.WithPlugin(targetloadpacking.Name, targetloadpacking.NewB200),
		app.WithPlugin(loadvariationriskbalancing.Name, loadvariationriskbalancing.NewB200),
		app.WithPlugin(noderesourcetopology.Name, noderesourcetopology.NewB200),
		app.WithPlugin(networkoverhead.Name, networkoverhead.NewB200),
	}
}
